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
  EF_GENERIC = 1 << 20
};

SC_DEVINL float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x*sigmoid(1.702x) with one MUFU: sigmoid(y) = 0.5 + 0.5 tanh(y/2)
SC_DEVINL float qgelu_fast(float x) { return x * fmaf(0.5f, tanh_fast(0.851f * x), 0.5f); }
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
    if constexpr ((F & EF_RESID) != 0) {
      v.x += pre.x; v.y += pre.y; v.z += pre.z; v.w += pre.w;    // prefetched residual
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

// host helpers (gemm_tc.cu)
int sc_get_tensor_map(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride_elems, uint32_t box0, uint32_t box1,
                      CUtensorMap* out);
int sc_select_epilogue(const sc_gemm_desc* d, int splits);
