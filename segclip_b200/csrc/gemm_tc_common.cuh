// Shared device helpers of the tcgen05 GEMM kernels (1-CTA and 2-CTA variants): PTX wrappers, smem descriptors,
// compile-time specialised epilogues.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace tc {


constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (NUM_EPI_WARPS + 2) * 32;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
SC_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SC_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SC_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
SC_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
SC_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 28)) {  // a protocol bug: fail loudly instead of hanging the GPU
      printf("segclip_b200 gemm_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
SC_DEVINL void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
SC_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SC_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SC_DEVINL void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
SC_DEVINL void tcgen05_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// ---- warp-uniform issue: the whole warp runs the producer / MMA loop (addresses and descriptors stay in uniform
// registers) and one elected lane executes the instruction.  With `if (lane == 0) { loop }` ptxas cannot prove
// uniformity and wraps every UTCHMMA in ~20 instructions of R2UR.BROADCAST / ELECT / BRA.U.ANY, which is slower than
// the N = 64 MMAs themselves.  elect.sync with a full mask always names the same lane, so tcgen05.commit still tracks
// the MMAs issued before it.
SC_DEVINL void tcgen05_mma_f16_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
SC_DEVINL void tcgen05_mma_f16_ts_e(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
SC_DEVINL void tcgen05_commit_e(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
SC_DEVINL void mbar_expect_tx_e(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
      : "memory");
}
SC_DEVINL void tma_load_2d_e(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
SC_DEVINL void tma_load_3d_e(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
SC_DEVINL void tma_prefetch_2d_e(const CUtensorMap* map, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];\n\t}" ::"l"((uint64_t)map), "r"(c0), "r"(c1)
      : "memory");
}
SC_DEVINL void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = (uint32_t*)v;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- compile-time specialised epilogues for the hot-path GEMMs (everything else: EF_GENERIC runtime path) ----
enum : int {
  EF_BIAS = 1,          // + bias[n]
  EF_QGELU = 2,         // QuickGELU
  EF_C2 = 4,            // also store the pre-activation (bf16)
  EF_RESID = 8,         // + residual (fp32)
  EF_OUT_F32 = 16,      // fp32 output (default bf16)
  EF_MULAUX_QGELU = 32, // * QuickGELU'(aux) (fused activation backward, aux bf16)
  EF_ATOMIC = 64,       // split-K: atomic accumulate into fp32 C
  EF_GELU = 128,        // exact erf GELU (MAE decoder / proj_o MLP)
  EF_MULAUX_GELU = 256, // * GELU'(aux)
  EF_ACCUM = 512,       // C += result (fp32 read-modify-write, single split; old C prefetched like a residual)
  EF_RESID_BF = 1024,   // + residual (bf16), bf16 output: the bf16 residual stream (2-CTA kernel: residual tile through the TMA)
  EF_ROWDOT = 2048,     // plain bf16 output + per-head row dots with a second bf16 tile (attention delta from the out_proj dgrad)
  EF_C2_DERIV = 4096,   // with EF_C2: the second output is act'(v), not v (the sigmoid / erf is at hand in the forward epilogue)
  EF_MULAUX_DERIV = 8192, // * aux, aux = act'(pre-activation) stored by the forward (3 instructions per element instead of ~10)
  EF_GENERIC = 1 << 20
};

SC_DEVINL float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x*sigmoid(1.702x) with one MUFU: sigmoid(y) = 0.5 + 0.5 tanh(y/2)
SC_DEVINL float qgelu_fast(float x) { return x * fmaf(0.5f, tanh_fast(0.851f * x), 0.5f); }
// y = QuickGELU(x), g = QuickGELU'(x) from ONE tanh: g = s (1 + 1.702 x (1 - s)) = s + 1.702 y (1 - s)
SC_DEVINL void qgelu_both(float x, float& y, float& g) {
  const float s = fmaf(0.5f, tanh_fast(0.851f * x), 0.5f);
  y = x * s;
  g = fmaf(1.702f * y, 1.0f - s, s);
}
template <int EF>
SC_DEVINL void act_both(float x, float& y, float& g) {
  if constexpr ((EF & EF_QGELU) != 0) qgelu_both(x, y, g);
  else { y = act_fwd(x, SC_ACT_GELU_ERF); g = act_grad(x, SC_ACT_GELU_ERF); }
}
SC_DEVINL float qgelu_grad_fast(float x) {
  const float s = fmaf(0.5f, tanh_fast(0.851f * x), 0.5f);
  return s * fmaf(1.702f * x, 1.0f - s, 1.0f);
}
SC_DEVINL void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
SC_DEVINL float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
SC_DEVINL uint2 pack4_bf16(const float4& v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *(uint32_t*)&a;
  u.y = *(uint32_t*)&b;
  return u;
}

template <int F>
SC_DEVINL void epi4_fast(const EpiParams& p, int m, int n, float4 v, const float4& b4, const float4& pre) {
  if constexpr (F == EF_GENERIC) {
    epi_store4(p, m, n, v);
  } else {
    const long off = (long)m * p.ldc + n;
    if constexpr ((F & EF_BIAS) != 0) { v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w; }
    if constexpr ((F & EF_C2) != 0) *(uint2*)((bf16*)p.C2 + off) = pack4_bf16(v);
    if constexpr ((F & EF_QGELU) != 0) { v.x = qgelu_fast(v.x); v.y = qgelu_fast(v.y); v.z = qgelu_fast(v.z); v.w = qgelu_fast(v.w); }
    if constexpr ((F & EF_GELU) != 0) {
      v.x = act_fwd(v.x, SC_ACT_GELU_ERF); v.y = act_fwd(v.y, SC_ACT_GELU_ERF);
      v.z = act_fwd(v.z, SC_ACT_GELU_ERF); v.w = act_fwd(v.w, SC_ACT_GELU_ERF);
    }
    if constexpr ((F & EF_MULAUX_QGELU) != 0) {
      const uint32_t ux = __float_as_uint(pre.x), uy = __float_as_uint(pre.y);   // prefetched bf16x4
      const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&ux), b = __bfloat1622float2(*(const __nv_bfloat162*)&uy);
      v.x *= qgelu_grad_fast(a.x); v.y *= qgelu_grad_fast(a.y); v.z *= qgelu_grad_fast(b.x); v.w *= qgelu_grad_fast(b.y);
    }
    if constexpr ((F & EF_MULAUX_GELU) != 0) {
      const uint32_t ux = __float_as_uint(pre.x), uy = __float_as_uint(pre.y);
      const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&ux), b = __bfloat1622float2(*(const __nv_bfloat162*)&uy);
      v.x *= act_grad(a.x, SC_ACT_GELU_ERF); v.y *= act_grad(a.y, SC_ACT_GELU_ERF);
      v.z *= act_grad(b.x, SC_ACT_GELU_ERF); v.w *= act_grad(b.y, SC_ACT_GELU_ERF);
    }
    if constexpr ((F & (EF_RESID | EF_ACCUM)) != 0) {
      v.x += pre.x; v.y += pre.y; v.z += pre.z; v.w += pre.w;    // prefetched residual / old C
    }
    if constexpr ((F & EF_ATOMIC) != 0) {
      // one 16-byte vector reduction instead of four scalar atomics (split-K wgrad: fp32 accumulate in L2)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"((float*)p.C + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    } else if constexpr ((F & EF_OUT_F32) != 0) {
      *(float4*)((float*)p.C + off) = v;
    } else {
      *(uint2*)((bf16*)p.C + off) = pack4_bf16(v);
    }
  }
}

SC_DEVINL void sts128b(uint32_t addr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
SC_DEVINL uint4 lds128b(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
  return u;
}
SC_DEVINL uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}
SC_DEVINL float2 unpack2_bf16(uint32_t u) { return __bfloat1622float2(*(const __nv_bfloat162*)&u); }

// ---- per-chunk epilogue driver (one warp, 32 rows x 32 accumulator columns) --------------------------------------
// bf16-output kinds do all column-wise math in the row-per-thread domain, pack to bf16 and transpose only 64 B per row
// through the warp's smem patch (half the shared-memory traffic of an fp32 transpose: with K = 768 the UMMA operand
// fetch already uses ~70 % of the SM's 128 B/clk shared-memory bandwidth, so the epilogue's share decides the speed).
// fp32-output kinds (residual add, split-K accumulate, generic) transpose fp32.
template <int EF>
struct EpiKind {
  static constexpr bool bf16_path = (EF != EF_GENERIC) && (EF & (EF_OUT_F32 | EF_ATOMIC | EF_RESID)) == 0;
  static constexpr bool aux = (EF != EF_GENERIC) && (EF & (EF_MULAUX_QGELU | EF_MULAUX_GELU | EF_MULAUX_DERIV)) != 0;
};

template <int EF>
SC_DEVINL void epi_prefetch(const EpiParams& ep, int lane, int mrow0, int n0, float4& b4, float4 (&pre)[8]) {
  if constexpr (EpiKind<EF>::bf16_path) {
    if constexpr (EpiKind<EF>::aux) {
      const int c = lane & 3, n = n0 + c * 8;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int m = mrow0 + 8 * t + (lane >> 2);
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (m < ep.M && n < ep.N) u = *(const uint4*)((const bf16*)ep.mul_aux + (long)m * ep.ldc + n);
        pre[t] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
      }
    }
  } else {
    const int l7 = lane & 7, l3 = lane >> 3;
    const int n = n0 + l7 * 4;
    b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < ep.N) {
      if constexpr (EF != EF_GENERIC && (EF & EF_BIAS) != 0) b4 = __ldg((const float4*)(ep.bias + n));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = mrow0 + 4 * i + l3;
        pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (EF != EF_GENERIC && (EF & EF_RESID) != 0) {
          if (m < ep.M) pre[i] = *(const float4*)((const float*)ep.residual + (long)m * ep.ldr + n);
        }
        if constexpr (EF != EF_GENERIC && (EF & EF_ACCUM) != 0) {
          if (m < ep.M) pre[i] = *(const float4*)((const float*)ep.C + (long)m * ep.ldc + n);
        }
      }
    }
  }
}

template <int EF>
SC_DEVINL void epi_finish(const EpiParams& ep, float (&v)[32], uint32_t stage, int lane, int mrow0, int n0, const float4& b4,
                          const float4 (&pre)[8]) {
  if constexpr (EpiKind<EF>::bf16_path) {
    // ---- row-per-thread math
    if constexpr ((EF & EF_BIAS) != 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0 + 4 * j < ep.N) {                       // warp-uniform
          const float4 b = __ldg((const float4*)(ep.bias + n0 + 4 * j));
          v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      }
    }
    const uint32_t rowaddr = stage + lane * 64;
    const int sw = (lane >> 1) & 3;
    __syncwarp();
    if constexpr ((EF & EF_C2) != 0 && (EF & EF_C2_DERIV) != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t gp[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float g0, g1;
          act_both<EF>(v[8 * j + 2 * k], v[8 * j + 2 * k], g0);
          act_both<EF>(v[8 * j + 2 * k + 1], v[8 * j + 2 * k + 1], g1);
          gp[k] = pack2_bf16(g0, g1);
        }
        sts128b(rowaddr + ((j ^ sw) << 4), make_uint4(gp[0], gp[1], gp[2], gp[3]));
      }
    } else {
    if constexpr ((EF & EF_C2) != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128b(rowaddr + ((j ^ sw) << 4), make_uint4(pack2_bf16(v[8 * j], v[8 * j + 1]), pack2_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                       pack2_bf16(v[8 * j + 4], v[8 * j + 5]), pack2_bf16(v[8 * j + 6], v[8 * j + 7])));
    }
    if constexpr ((EF & EF_QGELU) != 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = qgelu_fast(v[j]);
    }
    if constexpr ((EF & EF_GELU) != 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], SC_ACT_GELU_ERF);
    }
    }
    constexpr uint32_t out_tile = (EF & EF_C2) != 0 ? 2048u : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128b(rowaddr + out_tile + ((j ^ sw) << 4), make_uint4(pack2_bf16(v[8 * j], v[8 * j + 1]), pack2_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                                 pack2_bf16(v[8 * j + 4], v[8 * j + 5]), pack2_bf16(v[8 * j + 6], v[8 * j + 7])));
    __syncwarp();
    // ---- transposed, coalesced stores: 4 lanes x 16 B cover the 64 B of one row, 8 rows per instruction
    const int c = lane & 3, n = n0 + c * 8;
    float cs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int row = 8 * t + (lane >> 2);
      const int m = mrow0 + row;
      const uint32_t a = stage + row * 64 + ((c ^ ((row >> 1) & 3)) << 4);
      if (m < ep.M && n < ep.N) {
        const long off = (long)m * ep.ldc + n;
        if constexpr ((EF & EF_C2) != 0) *(uint4*)((bf16*)ep.C2 + off) = lds128b(a);
        uint4 u = lds128b(a + out_tile);
        if constexpr (EpiKind<EF>::aux) {
          const uint32_t aw[4] = {__float_as_uint(pre[t].x), __float_as_uint(pre[t].y), __float_as_uint(pre[t].z), __float_as_uint(pre[t].w)};
          uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 x = unpack2_bf16(uw[k]), g = unpack2_bf16(aw[k]);
            float g0, g1;
            if constexpr ((EF & EF_MULAUX_DERIV) != 0) { g0 = g.x; g1 = g.y; }
            else if constexpr ((EF & EF_MULAUX_QGELU) != 0) { g0 = qgelu_grad_fast(g.x); g1 = qgelu_grad_fast(g.y); }
            else { g0 = act_grad(g.x, SC_ACT_GELU_ERF); g1 = act_grad(g.y, SC_ACT_GELU_ERF); }
            uw[k] = pack2_bf16(x.x * g0, x.y * g1);
          }
          u = make_uint4(uw[0], uw[1], uw[2], uw[3]);
        }
        *(uint4*)((bf16*)ep.C + off) = u;
        if (ep.colsum_out) {                          // uniform branch: column sums of what was stored
          const uint32_t uw2[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 x = unpack2_bf16(uw2[k]);
            cs[2 * k] += x.x;
            cs[2 * k + 1] += x.y;
          }
        }
      }
    }
    if (ep.colsum_out) {
      // reduce over the 8 row groups of the warp (lanes with equal lane & 3), then one vector reduction per 4 columns
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4);
        cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
        cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
      }
      if (lane < 4 && n < ep.N) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.colsum_out + n), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.colsum_out + n + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
      }
    }
  } else {
    const int l7 = lane & 7, l3 = lane >> 3;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(stage + lane * 128 + ((j ^ l7) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int n = n0 + l7 * 4;
    if (n < ep.N) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + l3;
        const float4 x = lds128(stage + r * 128 + ((l7 ^ (r & 7)) << 4));
        const int m = mrow0 + r;
        if (m < ep.M) epi4_fast<EF>(ep, m, n, x, b4, pre[i]);
      }
    }
  }
}

// ---- TMA-store flavour of the bf16 epilogue (2-CTA kernel) --------------------------------------------------------
// ncu on the c_fc forward showed the L1TEX data pipe -- shared by UMMA operand fetch, LDS/STS and global LD/ST
// wavefronts -- 100 % busy: per 128 x 256 CTA tile 3072 (UMMA) + 3684 (smem transpose) + 1273 (global) wavefronts against
// 6144 cycles of MMA.  Here the row-per-thread values are packed to bf16, written once into a 32 x 32 box in the
// SWIZZLE_64B layout (conflict-free 16-byte stores) and handed to the TMA (cp.async.bulk.tensor store): no transposed
// read-back, no per-lane global stores.  Four 2 KB boxes per warp rotate; a box is rewritten only after the bulk group
// that read it has completed (cp.async.bulk.wait_group.read).
// (whole warp calls these; the elected lane -- always the same one for a full mask -- owns the bulk groups)
SC_DEVINL void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n\t}"
      ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
SC_DEVINL void bulk_commit() {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e cp.async.bulk.commit_group;\n\t}" ::: "memory");
}
template <int N>
SC_DEVINL void bulk_wait_read() {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e cp.async.bulk.wait_group.read %0;\n\t}" ::"n"(N) : "memory");
}
SC_DEVINL void bulk_wait_all() {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e cp.async.bulk.wait_group 0;\n\t}" ::: "memory");
}

template <int EF>
struct EpiTma {
  static constexpr bool value = EpiKind<EF>::bf16_path && !EpiKind<EF>::aux && (EF == EF_GENERIC || (EF & EF_RESID_BF) == 0);
  static constexpr bool two = value && (EF & EF_C2) != 0;
};

// ---- bf16 epilogues with a second bf16 INPUT tile (residual add, activation-gradient multiply), 2-CTA kernel ----------
// The input tile of a chunk is TMA-loaded into the very 32 x 32 SWIZZLE_64B box its result is stored from: every lane
// reads its own row (4 conflict-free 16-byte loads), combines it with the accumulator row it holds, writes the result back
// over it and the box leaves through the TMA.  The LSU version of this (`epi_prefetch` + transposed `epi_finish`) read the
// operand with per-lane global loads -- ncu: L1TEX data pipe saturated, activation-gradient dgrad at 63 % of the tensor
// peak, fp32-residual out_proj at 47 %.  A warp's four boxes are its four chunks of a tile: all four loads are issued at
// tile start (while the tile's main loop still runs) on one mbarrier each.
template <int EF>
struct EpiAuxTma {
  static constexpr bool value = (EF != EF_GENERIC) && (EF & (EF_RESID_BF | EF_MULAUX_QGELU | EF_MULAUX_GELU | EF_MULAUX_DERIV | EF_ROWDOT)) != 0 &&
                                (EF & (EF_OUT_F32 | EF_ATOMIC | EF_RESID | EF_C2)) == 0;
};

template <int EF>
SC_DEVINL void epi_finish_aux_tma(const EpiParams& ep, float (&v)[32], uint32_t box, int lane, int mrow0, int n0, int c,
                                  const float4& breg, const CUtensorMap* tmC, float& dotacc) {
  if constexpr ((EF & EF_BIAS) != 0) {
    const int src = c * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] += __shfl_sync(0xffffffffu, breg.x, src + j);
      v[4 * j + 1] += __shfl_sync(0xffffffffu, breg.y, src + j);
      v[4 * j + 2] += __shfl_sync(0xffffffffu, breg.z, src + j);
      v[4 * j + 3] += __shfl_sync(0xffffffffu, breg.w, src + j);
    }
  }
  const int sw = (lane >> 1) & 3;
  const uint32_t row = box + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = row + ((j ^ sw) << 4);
    const uint4 u = lds128b(a);
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = unpack2_bf16(uw[k]);
      float o0 = v[8 * j + 2 * k], o1 = v[8 * j + 2 * k + 1];
      if constexpr ((EF & EF_RESID_BF) != 0) { o0 += x.x; o1 += x.y; }
      if constexpr ((EF & EF_MULAUX_QGELU) != 0) { o0 *= qgelu_grad_fast(x.x); o1 *= qgelu_grad_fast(x.y); }
      if constexpr ((EF & EF_MULAUX_GELU) != 0) { o0 *= act_grad(x.x, SC_ACT_GELU_ERF); o1 *= act_grad(x.y, SC_ACT_GELU_ERF); }
      if constexpr ((EF & EF_MULAUX_DERIV) != 0) { o0 *= x.x; o1 *= x.y; }
      if constexpr ((EF & EF_ROWDOT) != 0) dotacc = fmaf(o0, x.x, fmaf(o1, x.y, dotacc));
      ow[k] = pack2_bf16(o0, o1);
    }
    sts128b(a, make_uint4(ow[0], ow[1], ow[2], ow[3]));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the TMA
  __syncwarp();
  tma_store_2d(tmC, box, n0, mrow0);
  bulk_commit();
  if constexpr ((EF & EF_ROWDOT) != 0) {
    // chunks come in pairs per 64-column head (tile column offsets are multiples of 128): the odd chunk completes the dot
    if (c & 1) {
      const int m = mrow0 + lane;
      if (m < ep.M) {
        const int b = m / ep.dot_L, l = m - b * ep.dot_L;
        ep.dot_out[((long)b * (ep.N >> 6) + (n0 >> 6)) * ep.dot_L + l] = dotacc;
      }
      dotacc = 0.f;
    }
  }
  if (ep.colsum_out) {                                 // uniform branch: column sums of what was stored (bias gradient)
    // rows >= M and columns >= N hold exact zeros (zero-filled operands and input tile), so no row mask is needed
    const int c4 = lane & 3, n = n0 + c4 * 8;
    float cs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int r = 8 * t + (lane >> 2);
      const uint4 u = lds128b(box + r * 64 + ((c4 ^ ((r >> 1) & 3)) << 4));
      const uint32_t uw2[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = unpack2_bf16(uw2[k]);
        cs[2 * k] += x.x;
        cs[2 * k + 1] += x.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4);
      cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
      cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
    }
    if (lane < 4 && n < ep.N) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.colsum_out + n), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.colsum_out + n + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
    }
  }
}

// ---- fp32 residual stream (out_proj / c_proj forward): C(fp32) = acc + bias + residual(fp32), both through the TMA ----
// Same idea with 32 x 32 fp32 boxes (4 KB, SWIZZLE_128B): a warp's 8 KB of staging hold two, so the loads of chunks 2 and
// 3 are issued as soon as the stores of chunks 0 and 1 have read their box.
template <int EF>
struct EpiResTma {
  static constexpr bool value = (EF != EF_GENERIC) && (EF & EF_RESID) != 0 && (EF & EF_OUT_F32) != 0 &&
                                (EF & (EF_ATOMIC | EF_ACCUM | EF_C2 | EF_MULAUX_QGELU | EF_MULAUX_GELU | EF_QGELU | EF_GELU)) == 0;
};

template <int EF>
SC_DEVINL void epi_finish_res_tma(const EpiParams& ep, float (&v)[32], uint32_t box, int lane, int mrow0, int n0, int c,
                                  const float4& breg, const CUtensorMap* tmC) {
  if constexpr ((EF & EF_BIAS) != 0) {
    const int src = c * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] += __shfl_sync(0xffffffffu, breg.x, src + j);
      v[4 * j + 1] += __shfl_sync(0xffffffffu, breg.y, src + j);
      v[4 * j + 2] += __shfl_sync(0xffffffffu, breg.z, src + j);
      v[4 * j + 3] += __shfl_sync(0xffffffffu, breg.w, src + j);
    }
  }
  const uint32_t row = box + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = row + ((j ^ sw) << 4);
    const float4 r = lds128(a);
    sts128(a, v[4 * j] + r.x, v[4 * j + 1] + r.y, v[4 * j + 2] + r.z, v[4 * j + 3] + r.w);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  tma_store_2d(tmC, box, n0, mrow0);
  bulk_commit();
}

// one warp, one chunk of 32 rows x 32 accumulator columns; `g` = running chunk counter of the warp (buffer rotation)
template <int EF>
SC_DEVINL void epi_finish_tma(const EpiParams& ep, float (&v)[32], uint32_t stage, int lane, int mrow0, int n0, uint32_t g,
                              int c, const float4& breg, const CUtensorMap* tmC, const CUtensorMap* tmC2) {
  constexpr bool two = EpiTma<EF>::two;
  const uint32_t buf_out = stage + ((two ? 2 * g + 1 : g) & 3u) * 2048u;
  const uint32_t buf_c2 = stage + ((2 * g) & 3u) * 2048u;
  bulk_wait_read<two ? 1 : 3>();                     // the group that last read these boxes has retired
  if constexpr ((EF & EF_BIAS) != 0) {
    // breg = the 4 bias values of columns 4*lane.. of this warp's 128-column range, loaded once per tile BEFORE the wait
    // for the accumulator (the L1 left beside 224 KB of smem is ~4 KB: a per-chunk __ldg exposed an L2 round trip per chunk
    // -- the top stall of the old epilogue in the ncu source view)
    const int src = c * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] += __shfl_sync(0xffffffffu, breg.x, src + j);
      v[4 * j + 1] += __shfl_sync(0xffffffffu, breg.y, src + j);
      v[4 * j + 2] += __shfl_sync(0xffffffffu, breg.z, src + j);
      v[4 * j + 3] += __shfl_sync(0xffffffffu, breg.w, src + j);
    }
  }
  const int sw = (lane >> 1) & 3;
  __syncwarp();
  if constexpr (two && (EF & EF_C2_DERIV) != 0) {
    // second output = act'(v): activation and derivative from the same sigmoid / erf
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t gp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float g0, g1;
        act_both<EF>(v[8 * j + 2 * k], v[8 * j + 2 * k], g0);
        act_both<EF>(v[8 * j + 2 * k + 1], v[8 * j + 2 * k + 1], g1);
        gp[k] = pack2_bf16(g0, g1);
      }
      sts128b(buf_c2 + lane * 64 + ((j ^ sw) << 4), make_uint4(gp[0], gp[1], gp[2], gp[3]));
    }
  } else {
  if constexpr (two) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128b(buf_c2 + lane * 64 + ((j ^ sw) << 4), make_uint4(pack2_bf16(v[8 * j], v[8 * j + 1]), pack2_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                                pack2_bf16(v[8 * j + 4], v[8 * j + 5]), pack2_bf16(v[8 * j + 6], v[8 * j + 7])));
  }
  if constexpr ((EF & EF_QGELU) != 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = qgelu_fast(v[j]);
  }
  if constexpr ((EF & EF_GELU) != 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], SC_ACT_GELU_ERF);
  }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128b(buf_out + lane * 64 + ((j ^ sw) << 4), make_uint4(pack2_bf16(v[8 * j], v[8 * j + 1]), pack2_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                               pack2_bf16(v[8 * j + 4], v[8 * j + 5]), pack2_bf16(v[8 * j + 6], v[8 * j + 7])));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the TMA
  __syncwarp();
  if (n0 < ep.N && mrow0 < ep.M) {                   // warp-uniform
    if constexpr (two) tma_store_2d(tmC2, buf_c2, n0, mrow0);
    tma_store_2d(tmC, buf_out, n0, mrow0);
  }
  bulk_commit();
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B.
//   K-major : rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO); LBO unused (1).
//   MN-major: 64-element (128 B) MN chunks; k rows 128 B apart, 8-k-row atoms SBO=1024 B apart,
//             consecutive MN chunks LBO = 64 rows * 128 B = 8192 B apart (one TMA box each).
template <bool MN_MAJOR>
SC_DEVINL uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(MN_MAJOR ? (8192 >> 4) : 1) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}


}  // namespace tc

// Split-K factor for a persistent kernel with `workers` CTAs (or CTA pairs): every work item is one (tile, split) of
// kb_per k-blocks and the workers take items round-robin, so the run lasts waves * (kb_per + per-item overhead).  The old
// rule (fill two waves) left the last wave 20-55 % empty on the ViT-B wgrads (e.g. 36 tiles x 5 splits = 180 items on 74
// pairs = 3 waves for 2.4 waves of work); this picks the split count with the lowest modelled time, fewest splits on ties.
static inline int sc_pick_splits(int tiles, int kb_total, int workers) {
  int best = 1;
  long best_cost = -1;
  int smax = kb_total / 4 < 64 ? kb_total / 4 : 64;
  if (smax < 1) smax = 1;
  for (int s = 1; s <= smax; ++s) {
    const int kb_per = (kb_total + s - 1) / s;
    if ((kb_total + kb_per - 1) / kb_per != s) continue;      // same launch as a smaller s
    const long items = (long)tiles * s;
    const long waves = (items + workers - 1) / workers;
    const long cost = waves * (kb_per + 3);                    // ~3 k-blocks of prologue / atomic epilogue per item
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

// host helpers (gemm_tc.cu)
int sc_get_tensor_map(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride_elems, uint32_t box0, uint32_t box1,
                      CUtensorMap* out);
int sc_get_tensor_map_sw(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride_elems, uint32_t box0, uint32_t box1,
                         int swizzle_bytes, CUtensorMap* out);
int sc_get_tensor_map_any(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride_elems, uint32_t box0, uint32_t box1,
                          int swizzle_bytes, int elem_bytes, CUtensorMap* out);
int sc_get_tensor_map_3d(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1_elems,
                         uint64_t stride2_elems, uint32_t box0, uint32_t box1, int swizzle_bytes, CUtensorMap* out);
int sc_get_tensor_map_nd(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* boxd,
                         int swizzle_bytes, CUtensorMap* out);
int sc_select_epilogue(const sc_gemm_desc* d, int splits);
