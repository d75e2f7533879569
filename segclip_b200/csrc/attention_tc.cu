// tcgen05 / TMEM attention backward for self-attention with head dim 64 and L <= 256 (vision L=196/48, text L=77
// causal): replaces the autograd backward of nn.MultiheadAttention's bmm/softmax/bmm
// (modules/module_seg_vit.py:189, modules/module_clip_ttransformer.py:46).
//
// One CTA per (sample, head).  Q, K, V, dO of the head are TMA-loaded once into 128B-swizzled smem (256 rows each).
// Work is done in the TRANSPOSED domain per (key tile kt, query tile qt) of 128 x 128, so every product is a plain UMMA
// whose operands are either the TMA tiles or the two bf16 tiles written by the softmax warps:
//   S^T  = K_kt Q_qt^T          (A K-major, B K-major,  N = 128)  -> TMEM cols [0,128)
//   dP^T = V_kt dO_qt^T         (A K-major, B K-major,  N = 128)  -> TMEM cols [128,256)
//   softmax warps: P^T = exp2(S^T c - lse_q), dS^T = P^T (dP^T - delta_q)  -> smem tiles (bf16, K-major, 128 x 128)
//   dV_kt += P^T  dO_qt         (A K-major tile,  B = dO MN-major, N = 64) -> TMEM cols [256,320)
//   dK_kt += dS^T Q_qt          (A K-major tile,  B = Q  MN-major, N = 64) -> TMEM cols [320,384)
//   dQ_qt += dS   K_kt          (A = dS^T tile read MN-major, B = K MN-major, N = 64) -> TMEM cols [384,448) / [448,512)
// TMEM is used completely (512 columns).  Warps 0-7: softmax + epilogues (lane quarter = warp % 4, column half = warp / 4);
// warp 16: MMA issue, warp 17: TMA producer (one elected lane each).
#include <stdlib.h>

#include "gemm_tc_common.cuh"

namespace {
using namespace tc;

constexpr int HD = 64;
constexpr int TILE = 128;
constexpr int ROWS = 256;                      // padded rows per operand
constexpr int OPER_BYTES = ROWS * 128;         // 32 KB
constexpr int PT_BYTES = TILE * TILE * 2;      // 32 KB (two 64-wide k-blocks of 16 KB)
constexpr int SM_Q = 0, SM_K = OPER_BYTES, SM_V = 2 * OPER_BYTES, SM_DO = 3 * OPER_BYTES;
constexpr int SM_PT = 4 * OPER_BYTES, SM_DST = SM_PT + PT_BYTES;
constexpr int SM_EPI = SM_DST + PT_BYTES;      // 16 warps x 2 KB: 32 x 32 bf16 boxes (SWIZZLE_64B) of the dQ / dK / dV TMA stores
constexpr int SM_LSE = SM_EPI + 16 * 2048;     // float[256] lse*log2e, float[256] delta
constexpr int SM_BAR = SM_LSE + 2 * ROWS * 4;
constexpr int SMEM_TOTAL = SM_BAR + 128;
constexpr int NSOFT = 16;                      // softmax / epilogue warps: 4 per TMEM lane quarter x 32 columns
constexpr int TC_THREADS = (NSOFT + 2) * 32;   // + MMA-issue warp + TMA producer warp
constexpr float LOG2E_F = 1.4426950408889634f;

constexpr uint32_t TM_ST = 0, TM_DPT = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

// smem descriptor with explicit leading / stride byte offsets (SWIZZLE_128B)
SC_DEVINL uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
SC_DEVINL float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SC_DEVINL uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}
SC_DEVINL void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, long bs, long rs, int B, int H, int L,
                                  float* __restrict__ delta, int hd = HD) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)B * L * H) return;
  const int h = t % H;
  const long row = t / H;
  const int b = row / L, i = row % L;
  const long off = (long)b * bs + (long)i * rs + h * hd;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    if (c * 8 >= hd) break;
    const uint4 a = *(const uint4*)(o + off + c * 8), g = *(const uint4*)(d_o + off + c * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = __bfloat1622float2(*(const __nv_bfloat162*)&aw[k]), y = __bfloat1622float2(*(const __nv_bfloat162*)&gw[k]);
      s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
    }
  }
  delta[((long)b * H + h) * L + i] = s;
}

SC_DEVINL void tmem_ld32_nowait(uint32_t taddr, float* v) {
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
}
SC_DEVINL void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = (uint32_t*)v;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
SC_DEVINL void store16_bf16(bf16* dst, const float* v, float scale) {
  uint4 u, w;
  u.x = pack_bf16(v[0] * scale, v[1] * scale); u.y = pack_bf16(v[2] * scale, v[3] * scale);
  u.z = pack_bf16(v[4] * scale, v[5] * scale); u.w = pack_bf16(v[6] * scale, v[7] * scale);
  w.x = pack_bf16(v[8] * scale, v[9] * scale); w.y = pack_bf16(v[10] * scale, v[11] * scale);
  w.z = pack_bf16(v[12] * scale, v[13] * scale); w.w = pack_bf16(v[14] * scale, v[15] * scale);
  *(uint4*)dst = u;
  *(uint4*)(dst + 8) = w;
}

#ifdef SC_ATT_TRACE
// debug build only: cycle-stamped events of CTA 0 (control warp + softmax warp 0), read back by sc_debug_attn_trace
__device__ long long g_trace[2][512];
__device__ int g_trace_n[2];
#define TRACE(who, tag)                                                                     \
  do {                                                                                      \
    if (blockIdx.x == 0 && lane == 0 && g_trace_n[who] < 255) {                              \
      const int i_ = g_trace_n[who]++;                                                      \
      g_trace[who][2 * i_] = (tag);                                                         \
      g_trace[who][2 * i_ + 1] = clock64();                                                 \
    }                                                                                       \
  } while (0)
#else
#define TRACE(who, tag) do { } while (0)
#endif

template <bool CAUSAL>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmGQ, const __grid_constant__ CUtensorMap tmGK,
                   const __grid_constant__ CUtensorMap tmGV, sc_attn_bwd_desc gd, const float* __restrict__ delta) {
  const sc_attn_desc& a = gd.fwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + SM_BAR);
  uint64_t* bar_load = bars;          // TMA bytes of the first 128 rows of Q, K, V, dO landed (one completion per item)
  uint64_t* bar_load1 = bars + 7;     // ... of rows 128..255 (items with two tiles)
  uint64_t* bar_free0 = bars + 8;     // commit: MMAs that read the first row group have retired (one per item with >1 pair)
  uint64_t* bar_free1 = bars + 9;     // commit: every MMA of the item has retired (one per item)
  uint64_t* bar_s = bars + 1;         // S / dP ready (commit, one per pair)
  uint64_t* bar_p = bars + 2;         // P / dS tiles written, S / dP consumed (all softmax warps, one per pair)
  uint64_t* bar_done = bars + 3;      // dV/dK/dQ MMAs of the pair retired (commit): tiles reusable
  uint64_t* bar_dkv = bars + 4;       // dV/dK of a key tile final (commit)
  uint64_t* bar_dkv_free = bars + 5;  // epilogue read dV/dK (all softmax warps)
  uint64_t* bar_dq_free = bars + 6;   // epilogue read dQ (all softmax warps, one per item)
  uint32_t* tmem_slot = (uint32_t*)(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.Lq;
  const int ntile = (L + TILE - 1) / TILE;
  const int total = a.H * a.B;                       // work items: (sample, head), persistent CTAs stride over them
  const bool pad = a.hd != HD;                       // head dim 32 / 48 in zero-padded 64-wide tiles (MAE decoder: 48)

  // The operands of an item -- rows b*L .. (+ntile*128), columns h*64 .. (+64) -- arrive as two groups of 128-row
  // boxes: the first rows of Q/K/V/dO are last read by the second-to-last pair, so the next item's first group is
  // loaded under the last pair and only the second group waits for the item's final MMAs.
  // (called by the whole control warp: one elected lane issues)
  // rg = row group of the item (rows rg*128 ..), buf = which 16 KB half of each operand buffer receives it
  auto issue_group = [&](int w, int rg, int buf) {
    const int h = w % a.H, b = w / a.H;
    uint64_t* bar = buf ? bar_load1 : bar_load;
    mbar_expect_tx_e(bar, 4 * TILE * 128);
    if (!pad) {
      tma_load_2d_e(&tmQ, bar, smem + SM_Q + buf * 16384, h * HD, b * L + rg * TILE);
      tma_load_2d_e(&tmK, bar, smem + SM_K + buf * 16384, h * HD, b * L + rg * TILE);
      tma_load_2d_e(&tmV, bar, smem + SM_V + buf * 16384, h * HD, b * L + rg * TILE);
      tma_load_2d_e(&tmdO, bar, smem + SM_DO + buf * 16384, h * HD, b * L + rg * TILE);
    } else {          // head dim < 64: maps {hd, H, rows}, the 64-wide box is zero-filled past the head (full box bytes are counted)
      tma_load_3d_e(&tmQ, bar, smem + SM_Q + buf * 16384, 0, h, b * L + rg * TILE);
      tma_load_3d_e(&tmK, bar, smem + SM_K + buf * 16384, 0, h, b * L + rg * TILE);
      tma_load_3d_e(&tmV, bar, smem + SM_V + buf * 16384, 0, h, b * L + rg * TILE);
      tma_load_3d_e(&tmdO, bar, smem + SM_DO + buf * 16384, 0, h, b * L + rg * TILE);
    }
  };
  // Single-tile items (L <= 128: text, MAE pass) use the two halves as a double buffer instead: item i lives in half
  // i & 1, so the next item's operands land while this one computes.
  const bool dbuf = ntile == 1;
  if (threadIdx.x == NSOFT * 32) {
    mbar_init(bar_load, 1);
    mbar_init(bar_load1, 1);
    mbar_init(bar_free0, 1);
    mbar_init(bar_free1, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, NSOFT);
    mbar_init(bar_done, 1);
    mbar_init(bar_dkv, 1);
    mbar_init(bar_dkv_free, NSOFT);
    mbar_init(bar_dq_free, NSOFT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == NSOFT) {
    __syncwarp();
    // dependent launch (common.cuh): barriers and TMEM under the previous kernel's tail, global reads after pdl_wait
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    pdl_wait();
    issue_group(blockIdx.x, 0, 0);
    if (ntile > 1) issue_group(blockIdx.x, 1, 1);
    else if ((int)(blockIdx.x + gridDim.x) < total) issue_group(blockIdx.x + gridDim.x, 0, 1);
  } else {
    pdl_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_launch_dependents();      // only now: a successor CTA sharing this SM must find this CTA's TMEM already allocated
  const float c = a.scale * LOG2E_F;

  if (warp == NSOFT) {
    {   // warp-uniform control flow; single-lane instructions are elected inside the *_e wrappers
      constexpr uint32_t ID_S = make_idesc(128, false, false);   // S, dP : A K-major, B K-major
      constexpr uint32_t ID_TT = make_idesc(64, true, true);     // dV, dK: A = tile read MN-major, B MN-major
      constexpr uint32_t ID_NT = make_idesc(64, false, true);    // dQ    : A = tile K-major,       B MN-major
      uint32_t ob = 0;                               // byte offset of the item's half (double-buffered single-tile items)
      auto issue_sdp = [&](int kt, int qt) {
        // S^T = K_kt Q_qt^T and dP^T = V_kt dO_qt^T (their TMEM columns were released by bar_p of the previous pair)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t dq_ = desc_sw128(sbase + SM_Q + ob + qt * 16384 + kk * 32, 16, 1024);
          const uint64_t dk_ = desc_sw128(sbase + SM_K + ob + kt * 16384 + kk * 32, 16, 1024);
          tcgen05_mma_f16_e(tmem + TM_ST, dq_, dk_, ID_S, kk > 0);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t do_ = desc_sw128(sbase + SM_DO + ob + qt * 16384 + kk * 32, 16, 1024);
          const uint64_t dv_ = desc_sw128(sbase + SM_V + ob + kt * 16384 + kk * 32, 16, 1024);
          tcgen05_mma_f16_e(tmem + TM_DPT, do_, dv_, ID_S, kk > 0);
        }
        tcgen05_commit_e(bar_s);
      };
      uint32_t ph_p = 0, ph_free = 0, ph_dq = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int np = CAUSAL ? ntile * (ntile + 1) / 2 : ntile * ntile;   // pairs of this item
        int pi = 0;
        bool have1 = ntile == 1;                     // second operand group waited for
        TRACE(0, 1);
        if (dbuf) {
          ob = (it & 1) * 16384u;
          mbar_wait((it & 1) ? bar_load1 : bar_load, (it >> 1) & 1);
        } else {
          mbar_wait(bar_load, it & 1);
        }
        tcgen05_fence_after();
        TRACE(0, 2);
        issue_sdp(0, 0);
        TRACE(0, 3);
        for (int kt = 0; kt < ntile; ++kt) {
          const int q_first = CAUSAL ? kt : 0;       // query tiles entirely before the key tile see nothing
          for (int qt = q_first; qt < ntile; ++qt) {
            mbar_wait(bar_p, ph_p);
            ph_p ^= 1;
            tcgen05_fence_after();
            TRACE(0, 4);
            // the next pair's S / dP go first: the softmax warps work on them while this pair's dV / dK / dQ run
            if (!have1 && pi + 1 < np) {               // every pair after the first touches rows 128..255
              mbar_wait(bar_load1, it & 1);
              tcgen05_fence_after();
              have1 = true;
            }
            if (qt + 1 < ntile) issue_sdp(kt, qt + 1);
            else if (kt + 1 < ntile) issue_sdp(kt + 1, CAUSAL ? kt + 1 : 0);
            if (qt == q_first && (kt > 0 || it > 0)) {   // dV/dK accumulators of the previous key tile must have been read
              mbar_wait(bar_dkv_free, ph_free);
              ph_free ^= 1;
              tcgen05_fence_after();
            }
            if (kt == 0 && qt == q_first && it > 0) {    // ... and the previous item's dQ
              mbar_wait(bar_dq_free, ph_dq);
              ph_dq ^= 1;
              tcgen05_fence_after();
            }
            TRACE(0, 5);
#pragma unroll
            for (int kq = 0; kq < 8; ++kq) {           // contraction over the 128 queries of this tile (tile rows)
              // tile [query rows][key cols] read as MN-major A: 64-key chunks 16 KB apart (LBO), 8-query groups 1 KB (SBO)
              const uint64_t dpa = desc_sw128(sbase + SM_PT + kq * 2048, 16384, 1024);
              const uint64_t dsa = desc_sw128(sbase + SM_DST + kq * 2048, 16384, 1024);
              const uint64_t dob = desc_sw128(sbase + SM_DO + ob + (qt * TILE + kq * 16) * 128, 8192, 1024);
              const uint64_t dqb = desc_sw128(sbase + SM_Q + ob + (qt * TILE + kq * 16) * 128, 8192, 1024);
              tcgen05_mma_f16_e(tmem + TM_DV, dpa, dob, ID_TT, (qt > q_first || kq > 0));
              tcgen05_mma_f16_e(tmem + TM_DK, dsa, dqb, ID_TT, (qt > q_first || kq > 0));
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {           // contraction over the 128 keys of this tile (tile columns)
              const uint64_t dsa = desc_sw128(sbase + SM_DST + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
              const uint64_t dkb = desc_sw128(sbase + SM_K + ob + (kt * TILE + kk * 16) * 128, 8192, 1024);
              tcgen05_mma_f16_e(tmem + TM_DQ + qt * 64, dsa, dkb, ID_NT, (kt > 0 || kk > 0));
            }
            tcgen05_commit_e(bar_done);
            TRACE(0, 6);
            ++pi;
            // the last pair reads only rows 128..255: once everything up to the second-to-last pair has retired the
            // producer warp may overwrite rows 0..127 with the next item
            if (pi == np - 1) tcgen05_commit_e(bar_free0);
          }
          tcgen05_commit_e(bar_dkv);
        }
        tcgen05_commit_e((dbuf && !(it & 1)) ? bar_free0 : bar_free1);
        TRACE(0, 7);
      }
    }
  } else if (warp == NSOFT + 1) {
    // ---------------- TMA producer: a UTMALDG blocks its warp for ~800 cycles per 16 KB box while the TMA unit drains its
    // queue (measured with SC_ATT_TRACE), so the loads have their own warp and never delay MMA issue
    const int np = CAUSAL ? ntile * (ntile + 1) / 2 : ntile * ntile;
    int it = 0;
    for (int w = blockIdx.x; w + (int)gridDim.x < total; w += gridDim.x, ++it) {
      const int wn = w + gridDim.x;
      if (!dbuf && !pad) {   // first touch of the next item goes to L2 now, a whole item ahead of the smem loads below
        const int h = wn % a.H, b = wn / a.H;
        for (int g = 0; g < ntile; ++g) {
          tma_prefetch_2d_e(&tmQ, h * HD, b * L + g * TILE);
          tma_prefetch_2d_e(&tmK, h * HD, b * L + g * TILE);
          tma_prefetch_2d_e(&tmV, h * HD, b * L + g * TILE);
          tma_prefetch_2d_e(&tmdO, h * HD, b * L + g * TILE);
        }
      }
      if (np > 1) {
        mbar_wait(bar_free0, it & 1);
        issue_group(wn, 0, 0);
        mbar_wait(bar_free1, it & 1);
        issue_group(wn, 1, 1);
      } else if (wn + (int)gridDim.x < total) {
        // double buffer: item it+2 goes where item `it` lives, once item `it` has retired (item it+1 was loaded earlier)
        mbar_wait((it & 1) ? bar_free1 : bar_free0, (it >> 1) & 1);
        issue_group(wn + gridDim.x, 0, it & 1);
      }
    }
  } else {
    // ---------------- softmax + epilogue warps: TMEM lane quarter = warp % 4, 32-column group = warp / 4
    const int quarter = warp & 3, cg = warp >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int r = quarter * 32 + lane;                  // row inside the tile
    uint32_t ph_s = 0, ph_done = 0, ph_dkv = 0;
    bool first_pair = true;                             // very first pair of this CTA: no earlier MMAs read the tiles
    // this thread's two query rows (one per query tile): lse and delta straight from global (coalesced); the next
    // item's values are requested before the epilogues of the current one, so their latency is never exposed
    float lse_c[2], dl_c[2];
    auto prefetch_rows = [&](int w) {                   // (registers are tight: holding the next item's values spilled them)
      const int h = w % a.H, b = w / a.H;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qi = t * TILE + r;
        const long o = ((long)b * a.H + h) * L + qi;
        if (qi < L && (lane & 7) == 0) {                // one request per 32-byte sector
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.lse + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(delta + o));
        }
      }
    };
    auto load_rows = [&](int w, float (&ls)[2], float (&dl)[2]) {
      const int h = w % a.H, b = w / a.H;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qi = t * TILE + r;
        const long o = ((long)b * a.H + h) * L + qi;
        ls[t] = qi < L ? a.lse[o] : 0.f;
        dl[t] = qi < L ? delta[o] : 0.f;
      }
    };
    // dV / dK / dQ leave through the TMA: 32 x 32 boxes per warp (lane quarter x one 32-column half of one tensor), packed
    // to bf16, written once into a SWIZZLE_64B staging box and stored with cp.async.bulk.tensor over a 3-D map {H*64, L, B}
    // whose second dimension clips the rows past the sample.  (The first version stored 32-byte pieces per lane: 32
    // distinct lines per store instruction; the LSU wavefronts of all 16 warps cost ~2 k cycles per key tile and ~5 k at
    // item ends -- on the softmax warps' critical path, SC_ATT_TRACE.)
    const uint32_t estage = sbase + SM_EPI + warp * 2048;
    // (16 columns at a time: the kernel runs at its 96-register cap -- 18 warps, five per SM sub-partition)
    auto stage_and_store = [&](const CUtensorMap* map, uint32_t tcol, bool live, float scale, int col0, int row0, int b) {
      bulk_wait_read<0>();                               // (elected lane) the previous store has read the staging box
      __syncwarp();
      const int sw = (lane >> 1) & 3;
      if (live) {
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          float v[16];
          tmem_ld16(tmem + lane_off + tcol + hf * 16, v);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            sts128u(estage + lane * 64 + (((hf * 2 + j) ^ sw) << 4), pack_bf16(v[8 * j] * scale, v[8 * j + 1] * scale),
                    pack_bf16(v[8 * j + 2] * scale, v[8 * j + 3] * scale), pack_bf16(v[8 * j + 4] * scale, v[8 * j + 5] * scale),
                    pack_bf16(v[8 * j + 6] * scale, v[8 * j + 7] * scale));
        }
      }
      return live;
    };
    // bias gradient of the q / k / v projection = column sums of dQ / dK / dV over all rows: taken from the staged bf16 box
    // (rows past L hold exact zeros: their P / dS entries are masked), transposed read-back like the GEMM epilogue's
    auto box_colsum = [&](float* out, bool live, int colh, int h) {
      if (out == nullptr || !live) return;              // warp-uniform
      const int col0 = h * a.hd + colh;
      __syncwarp();
      const int c4 = lane & 3;
      float cs[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int rr = 8 * t + (lane >> 2);
        const uint4 u = lds128b(estage + rr * 64 + ((c4 ^ ((rr >> 1) & 3)) << 4));
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 x = unpack2_bf16(uw[k]);
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
      if (lane < 4 && colh + c4 * 8 < a.hd) {             // (columns past a padded head hold zeros: skipped)
        float* o = out + col0 + c4 * 8;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
      }
    };
    auto release_and_store = [&](const CUtensorMap* map, bool live, int colh, int h, int row0, int b) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (live && !pad) {                                 // warp-uniform; rows past L inside the box are clipped by the map
        asm volatile(
            "{\n\t.reg .pred e;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "@e cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n\t}"
            ::"l"((uint64_t)map), "r"(estage), "r"(h * HD + colh), "r"(row0), "r"(b) : "memory");
      } else if (live) {                                  // padded head: map {hd, H, L, B}, columns past the head clipped too
        asm volatile(
            "{\n\t.reg .pred e;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "@e cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n\t}"
            ::"l"((uint64_t)map), "r"(estage), "r"(colh), "r"(h), "r"(row0), "r"(b) : "memory");
      }
      bulk_commit();
    };
    auto dkv_epilogue = [&](int kt, int h, int b) {
      // dV / dK of a key tile (TMEM lanes = keys): column groups 0,1 take the halves of dV, 2,3 those of dK
      mbar_wait(bar_dkv, ph_dkv);
      ph_dkv ^= 1;
      tcgen05_fence_after();
      const int row0 = kt * TILE + quarter * 32;
      const CUtensorMap* map = cg < 2 ? &tmGV : &tmGK;
      stage_and_store(map, (cg < 2 ? TM_DV : TM_DK) + (cg & 1) * 32, row0 < L, cg < 2 ? 1.0f : a.scale, 0, 0, 0);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dkv_free);
      release_and_store(map, row0 < L, (cg & 1) * 32, h, row0, b);
      box_colsum(cg < 2 ? gd.dv_colsum : gd.dk_colsum, row0 < L, (cg & 1) * 32, h);
    };
    auto dq_epilogue = [&](int h, int b) {
      // dQ (TMEM lanes = queries): bar_dkv of the item's last key tile was committed after every MMA of the item;
      // column groups 0,1 take the halves of query tile 0, groups 2,3 those of query tile 1
      const int qt = cg >> 1;
      const int row0 = qt * TILE + quarter * 32;
      const bool live = qt < ntile && row0 < L;
      stage_and_store(&tmGQ, TM_DQ + qt * 64 + (cg & 1) * 32, live, a.scale, 0, 0, 0);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dq_free);
      release_and_store(&tmGQ, live, (cg & 1) * 32, h, row0, b);
      box_colsum(gd.dq_colsum, live, (cg & 1) * 32, h);
    };
    // The read-out of an item's last dV / dK and of its dQ is deferred until this warp has delivered the first tile of
    // the NEXT item, so the MMA pipe never waits for the drain (the control warp holds the next item's accumulating
    // MMAs back until bar_dkv_free / bar_dq_free).
    int prev_h = -1, prev_b = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int h = w % a.H, b = w / a.H;
      load_rows(w, lse_c, dl_c);                         // L2 hits (prefetched below); consumed after the wait for S
      for (int kt = 0; kt < ntile; ++kt) {
        const int q_first = CAUSAL ? kt : 0;
        for (int qt = q_first; qt < ntile; ++qt) {
          const int qi = qt * TILE + r;                    // this thread's query
          const int key0 = kt * TILE + cg * 32;            // first key of this warp's column group
          const float lse2 = (qt ? lse_c[1] : lse_c[0]) * LOG2E_F, dl = qt ? dl_c[1] : dl_c[0];
          if (warp == 0) TRACE(1, 10);
          mbar_wait(bar_s, ph_s);
          ph_s ^= 1;
          tcgen05_fence_after();
          if (warp == 0) TRACE(1, 11);
          uint32_t pk[16], dk[16];                         // packed bf16 pairs of 32 columns
          const bool warp_live = (qt * TILE + quarter * 32 < L) && key0 < L && (!CAUSAL || key0 <= qt * TILE + quarter * 32 + 31);
          if (warp_live) {
            float s[32], dp[32];
            tmem_ld32_nowait(tmem + lane_off + TM_ST + cg * 32, s);
            tmem_ld32_nowait(tmem + lane_off + TM_DPT + cg * 32, dp);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (warp == 0) TRACE(1, 18);
            const int row_lo = qt * TILE + quarter * 32;       // the warp's first query
            const bool full = row_lo + 31 < L && key0 + 31 < L && (!CAUSAL || key0 + 31 <= row_lo);   // warp-uniform
            if (full) {                                        // interior block: no masks
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 = ex2f(fmaf(s[j], c, -lse2)), p1 = ex2f(fmaf(s[j + 1], c, -lse2));
                pk[j / 2] = pack_bf16(p0, p1);
                dk[j / 2] = pack_bf16(p0 * (dp[j] - dl), p1 * (dp[j + 1] - dl));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const int k0 = key0 + j;
                const bool ok0 = qi < L && k0 < L && (!CAUSAL || k0 <= qi);
                const bool ok1 = qi < L && k0 + 1 < L && (!CAUSAL || k0 + 1 <= qi);
                const float p0 = ok0 ? ex2f(fmaf(s[j], c, -lse2)) : 0.f;
                const float p1 = ok1 ? ex2f(fmaf(s[j + 1], c, -lse2)) : 0.f;
                const float d0 = ok0 ? p0 * (dp[j] - dl) : 0.f, d1 = ok1 ? p1 * (dp[j + 1] - dl) : 0.f;
                pk[j / 2] = pack_bf16(p0, p1);
                dk[j / 2] = pack_bf16(d0, d1);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = dk[j] = 0u;
          }
          if (warp == 0) TRACE(1, 12);
          if (!first_pair) {                               // the previous pair's MMAs must be done reading the tiles
            mbar_wait(bar_done, ph_done);
            ph_done ^= 1;
          }
          if (warp == 0) TRACE(1, 13);
          first_pair = false;
          // K-major 128B-swizzled tile [query rows][key cols]: k-block cg/2, 16-byte chunks (cg&1)*4 + j
          const uint32_t rowp = sbase + SM_PT + (cg >> 1) * 16384 + r * 128, rowd = sbase + SM_DST + (cg >> 1) * 16384 + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t off = (uint32_t)((((cg & 1) * 4 + j) ^ (r & 7)) << 4);
            sts128u(rowp + off, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            sts128u(rowd + off, dk[4 * j], dk[4 * j + 1], dk[4 * j + 2], dk[4 * j + 3]);
          }
          tcgen05_fence_before();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_p);
          if (warp == 0) TRACE(1, 14);
          // the previous key tile's dV / dK are read out only now: this pair's softmax went first, so the MMA pipe has
          // its next tiles while the accumulators are drained and stored
          if (qt == q_first && kt > 0) {
            dkv_epilogue(kt - 1, h, b);
            if (warp == 0) TRACE(1, 16);
          } else if (qt == q_first && prev_h >= 0) {
            dkv_epilogue(ntile - 1, prev_h, prev_b);
            dq_epilogue(prev_h, prev_b);
            if (warp == 0) TRACE(1, 17);
          }
        }
      }
      if (w + (int)gridDim.x < total) prefetch_rows(w + gridDim.x);
      prev_h = h;
      prev_b = b;
    }
    if (prev_h >= 0) {
      dkv_epilogue(ntile - 1, prev_h, prev_b);
      dq_epilogue(prev_h, prev_b);
    }
    bulk_wait_all();                                       // staging smem must outlive the last TMA stores
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == NSOFT) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05 / TMEM attention FORWARD (same scope: self-attention, head dim 64, L <= 256, optional causal mask): replaces
// the bmm / softmax / bmm of nn.MultiheadAttention (modules/module_seg_vit.py:189, module_clip_ttransformer.py:46).
//
// Persistent CTAs, 2 per SM (80 KB smem, 256 TMEM columns each), looping over work items (query tile of 128, head,
// sample): the next item's Q/K are TMA-loaded as soon as S has been computed and its V as soon as P V has retired, so
// loads run under the softmax / epilogue of the current item, and the two co-resident CTAs fill each other's bubbles.
// L <= 256 means a whole score row fits in TMEM: no online-softmax rescaling.
//   S = Q_tile K^T           (A, B from smem, K-major; N = L rounded up to 16)          -> TMEM cols [0, N)
//   softmax warps (thread = row): ONE pass, p = exp2(s c - shift) with a lagging row maximum as shift (see the kernel), row
//                            sum; P is written back as packed bf16
//                            INTO TMEM over the consumed S columns [0, N/2)  (tcgen05.st)
//   O = P V                  (A = P from TMEM, B = V from smem MN-major, N = 64)          -> TMEM cols [128, 192)
//   epilogue: O / rowsum -> bf16 -> global; lse = m scale + ln(rowsum)
// Warps 0-3: softmax + epilogue (TMEM lane quarter = warp); warp 4: TMA + MMA issue (one thread).
constexpr int F_SM_Q = 0, F_SM_K = TILE * 128, F_SM_V = F_SM_K + OPER_BYTES;
#ifndef SC_ATT_FWD_SOFT_WARPS
#define SC_ATT_FWD_SOFT_WARPS 4
#endif
// Optional 8-softmax-warp variant (-DSC_ATT_FWD_SOFT_WARPS=8): two warps per TMEM lane quarter, each taking half of the
// row's 32-column chunks (the 4-warp version runs at 2.4 warps per scheduler: issue slots 29 % busy, MUFU 33 %).
// Measured on B200: correct (same tests) but SLOWER, 0.165 vs 0.134 ms per vision layer -- the two named barriers, the
// smem detour and the lost TMEM-load prefetch cost more than the extra warps hide -- so 4 warps stay the default.
// The halves exchange row max / row sum through smem and a 64-thread named barrier; the upper half parks its packed P in shared memory until the lower half has read all of its S
// columns, because P (half as wide as S) lands on the lower half's columns.
constexpr int F_NSOFT = SC_ATT_FWD_SOFT_WARPS;          // 4 or 8
constexpr int F_SM_BAR = F_SM_V + OPER_BYTES;
constexpr int F_SM_X = F_SM_BAR + 64;                   // float [2 (max | sum)][2 (half)][128 rows]
constexpr int F_SM_P = F_SM_X + 2 * 2 * 128 * 4;        // parked P of the upper half: [3 chunks][128 rows][64 B], XOR-swizzled
constexpr int F_SM_O = ((F_SM_P + (F_NSOFT == 8 ? 3 * 128 * 64 : 0)) + 1023) / 1024 * 1024;   // 4 x 4 KB: O boxes of the TMA stores (SWIZZLE_128B)
constexpr int F_SMEM_TOTAL = F_SM_O + 4 * 4096;
constexpr int F_THREADS = (F_NSOFT + 1) * 32;
constexpr uint32_t F_TM_O = 128, F_TM_COLS = 256;

SC_DEVINL void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <bool CAUSAL>
__global__ void __launch_bounds__(F_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, sc_attn_desc a, int online) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + F_SM_BAR);
  uint64_t* bar_load = bars;       // TMA bytes landed
  uint64_t* bar_s = bars + 1;      // S ready (commit)
  uint64_t* bar_p = bars + 2;      // P written to TMEM (4 softmax warps)
  uint64_t* bar_o = bars + 3;      // O ready (commit)
  uint64_t* bar_v = bars + 4;      // V landed (needed only for P V: S and the softmax run under its flight)
  uint32_t* tmem_slot = (uint32_t*)(bars + 6);

  uint64_t* bar_free = bars + 5;   // O read out of TMEM by the 4 softmax warps: S / P / O columns reusable
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.Lq;
  const int ntile = (L + TILE - 1) / TILE;
  const int npad = (L + 15) & ~15;                 // MMA N of S, MMA K of P V
  const int total = ntile * a.H * a.B;
  const bool pad = a.hd != HD;                     // head dim 32 / 48 in zero-padded 64-wide tiles

  // (called by the whole control warp: one elected lane issues)
  auto issue_qk = [&](int w) {
    const int qt = w % ntile, h = (w / ntile) % a.H, b = w / (ntile * a.H);
    mbar_expect_tx_e(bar_load, TILE * 128 + ntile * TILE * 128);
    if (!pad) {
      tma_load_2d_e(&tmQ, bar_load, smem + F_SM_Q, h * HD, b * L + qt * TILE);
      tma_load_2d_e(&tmK, bar_load, smem + F_SM_K, h * HD, b * L);
    } else {
      tma_load_3d_e(&tmQ, bar_load, smem + F_SM_Q, 0, h, b * L + qt * TILE);
      tma_load_3d_e(&tmK, bar_load, smem + F_SM_K, 0, h, b * L);
    }
  };
  auto issue_v = [&](int w) {
    const int h = (w / ntile) % a.H, b = w / (ntile * a.H);
    mbar_expect_tx_e(bar_v, ntile * TILE * 128);
    if (!pad) tma_load_2d_e(&tmV, bar_v, smem + F_SM_V, h * HD, b * L);
    else tma_load_3d_e(&tmV, bar_v, smem + F_SM_V, 0, h, b * L);
  };

  if (threadIdx.x == F_NSOFT * 32) {
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, F_NSOFT);
    mbar_init(bar_o, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_free, F_NSOFT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == F_NSOFT) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(F_TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    pdl_wait();                      // (common.cuh) everything above ran under the previous kernel's tail
    issue_qk(blockIdx.x);
    issue_v(blockIdx.x);
  } else {
    pdl_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_launch_dependents();      // only now: a successor CTA sharing this SM must find this CTA's TMEM already allocated

  if (warp == F_NSOFT) {
    {   // warp-uniform control flow; single-lane instructions are elected inside the *_e wrappers
      const uint32_t id_s = make_idesc(npad, false, false);
      constexpr uint32_t ID_PV = make_idesc(64, false, true);
      uint32_t par = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, par ^= 1) {
        const int qt = w % ntile;
        const int wn = w + gridDim.x;
        TRACE(0, 20);
        mbar_wait(bar_load, par);
        tcgen05_fence_after();
        TRACE(0, 21);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t dq_ = desc_sw128(sbase + F_SM_Q + kk * 32, 16, 1024);
          const uint64_t dk_ = desc_sw128(sbase + F_SM_K + kk * 32, 16, 1024);
          tcgen05_mma_f16_e(tmem, dq_, dk_, id_s, kk > 0);
        }
        tcgen05_commit_e(bar_s);
        TRACE(0, 22);
        mbar_wait(bar_s, par);                       // Q / K consumed: the next item's may land
        TRACE(0, 23);
        if (wn < total) issue_qk(wn);
        mbar_wait(bar_v, par);
        mbar_wait(bar_p, par);
        tcgen05_fence_after();
        TRACE(0, 24);
        const int nk = CAUSAL ? min(npad, ((qt * TILE + TILE + 15) & ~15)) >> 4 : npad >> 4;   // keys past the tile's last query: P = 0
        for (int kk = 0; kk < nk; ++kk) {
          const uint64_t dv_ = desc_sw128(sbase + F_SM_V + kk * 16 * 128, 8192, 1024);
          tcgen05_mma_f16_ts_e(tmem + F_TM_O, tmem + kk * 8, dv_, ID_PV, kk > 0);
        }
        tcgen05_commit_e(bar_o);
        TRACE(0, 25);
        mbar_wait(bar_o, par);                       // V consumed
        TRACE(0, 26);
        if (wn < total) issue_v(wn);
        mbar_wait(bar_free, par);                    // O read out: the accumulator columns may be overwritten
        tcgen05_fence_after();
        TRACE(0, 27);
      }
    }
  } else {
    const int quarter = warp & 3, half = warp >> 2;       // half is always 0 with 4 softmax warps
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int r = quarter * 32 + lane;
    const float c = a.scale * LOG2E_F;
    float* xs = (float*)(smem + F_SM_X);
    uint32_t par = 0;
    if constexpr (F_NSOFT == 8) {
      const int nchunks = (npad + 31) >> 5;
      const int n1 = nchunks / 2 < 3 ? nchunks / 2 : 3;      // chunks of the upper half (its P waits in 24 KB of smem)
      const int c_lo = half ? nchunks - n1 : 0, c_hi = half ? nchunks : nchunks - n1;
      for (int w = blockIdx.x; w < total; w += gridDim.x, par ^= 1) {
        const int qt = w % ntile, h = (w / ntile) % a.H, b = w / (ntile * a.H);
        const int qi = qt * TILE + r;
        const bool warp_live = qt * TILE + quarter * 32 < L;   // same for both warps of the quarter
        float m = -INFINITY, sum = 0.f;
        mbar_wait(bar_s, par);
        tcgen05_fence_after();
        if (warp_live) {
          const int kmax = CAUSAL ? min(L, qi + 1) : L;
          float s[32];
          // ---- pass 1: max over this warp's chunks, then the row max across the two halves
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          for (int ch = c_lo; ch < c_hi; ++ch) {
            const int c0 = ch * 32;
            tmem_ld32_nowait(tmem + lane_off + c0, s);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c0 + 32 <= kmax) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], c0 + j < kmax ? s[j] : -INFINITY);
            }
          }
          m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          xs[half * 128 + r] = m;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
          m = fmaxf(m, xs[(half ^ 1) * 128 + r]);
          const float mc = (m == -INFINITY) ? 0.f : m * c;
          // ---- pass 2: p, partial row sum; the lower half stores P at once (its P columns lie inside its own, already
          // consumed S columns), the upper half parks it in shared memory
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
          const uint32_t prow = sbase + F_SM_P + r * 64;
          const int psw = (r >> 1) & 3;
          for (int ch = c_lo; ch < c_hi; ++ch) {
            const int i = ch - c_lo;
            {
              const int c0 = ch * 32;
              tmem_ld32_nowait(tmem + lane_off + c0, s);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              uint32_t q[16];
              if (c0 + 32 <= kmax) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float p0 = ex2f(fmaf(s[j], c, -mc)), p1 = ex2f(fmaf(s[j + 1], c, -mc));
                  s4[(j >> 1) & 3] += p0 + p1;
                  q[j >> 1] = pack_bf16(p0, p1);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float p0 = c0 + j < kmax ? ex2f(fmaf(s[j], c, -mc)) : 0.f;
                  const float p1 = c0 + j + 1 < kmax ? ex2f(fmaf(s[j + 1], c, -mc)) : 0.f;
                  s4[(j >> 1) & 3] += p0 + p1;
                  q[j >> 1] = pack_bf16(p0, p1);
                }
              }
              if (half == 0) {
                tmem_st16(tmem + lane_off + (c0 >> 1), q);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  sts128u(prow + i * 8192 + ((j ^ psw) << 4), q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
              }
            }
          }
          sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
          xs[256 + half * 128 + r] = sum;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");   // lower half has read all of its S; sums visible
          sum += xs[256 + (half ^ 1) * 128 + r];
          if (half == 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              if (i < n1) {
                uint32_t q[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 u = lds128b(prow + i * 8192 + ((j ^ psw) << 4));
                  q[4 * j] = u.x; q[4 * j + 1] = u.y; q[4 * j + 2] = u.z; q[4 * j + 3] = u.w;
                }
                tmem_st16(tmem + lane_off + ((c_lo + i) * 16), q);
              }
            }
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        mbar_wait(bar_o, par);
        tcgen05_fence_after();
        float o[32];
        if (warp_live) {
          tmem_ld32_nowait(tmem + lane_off + F_TM_O + half * 32, o);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free);
        if (warp_live && qi < L) {
          const float inv = 1.0f / sum;
          bf16* dst = (bf16*)a.o + (long)b * a.o_bs + (long)qi * a.o_rs + h * HD + half * 32;
          store16_bf16(dst, o, inv);
          store16_bf16(dst + 16, o + 16, inv);
          if (half == 0) a.lse[((long)b * a.H + h) * L + qi] = m * a.scale + logf(sum);
        }
      }
    } else
    for (int w = blockIdx.x; w < total; w += gridDim.x, par ^= 1) {
    const int qt = w % ntile, h = (w / ntile) % a.H, b = w / (ntile * a.H);
    const int qi = qt * TILE + r;
    const bool warp_live = qt * TILE + warp * 32 < L;
    float m = -INFINITY, sum = 0.f;
    if (warp == 0) TRACE(1, 30);
    mbar_wait(bar_s, par);
    tcgen05_fence_after();
    if (warp == 0) TRACE(1, 31);
    if (warp_live) {
      const int kmax = CAUSAL ? min(L, qi + 1) : L;          // this row sees keys [0, kmax)
      float bufa[32], bufb[32];                                // named buffers: static register indexing
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      if (online) {
        // ---- ONE pass over S (TMEM reads, ~64 B/clk/SM, are the floor of this kernel at head dim 64; two passes read every
        // score twice).  The exponent shift is a LAGGING row maximum: it starts as the maximum of the first 32 keys and is
        // raised only when a later chunk exceeds it by more than 2^8 in the exp2 domain; until then p = exp2(s c - shift)
        // may reach 256, which bf16 / the fp32 row sum hold exactly as well as values <= 1 (softmax is shift invariant,
        // O / rowsum and lse = shift + ln(rowsum) absorb it).  When the shift does move (rare; warp vote), the P chunks
        // already written are re-read from TMEM, scaled by exp2(old - new) and written back.
        float sh = -INFINITY;                                  // current shift, in units of s (not yet multiplied by c)
        auto online_chunk = [&](float (&s)[32], float (&nxt)[32], int c0) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 32 < npad) tmem_ld32_nowait(tmem + lane_off + c0 + 32, nxt);
          const bool edge = c0 + 32 > kmax;
          float cm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (!edge) {
#pragma unroll
            for (int j = 0; j < 32; ++j) cm4[j & 3] = fmaxf(cm4[j & 3], s[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              s[j] = c0 + j < kmax ? s[j] : -INFINITY;       // masked keys: exp2(-inf) = 0
              cm4[j & 3] = fmaxf(cm4[j & 3], s[j]);
            }
          }
          const float cm = fmaxf(fmaxf(cm4[0], cm4[1]), fmaxf(cm4[2], cm4[3]));
          const bool raise = (cm - sh) * c > 8.0f;             // also true for the first live chunk (sh = -inf)
          if (__any_sync(0xffffffffu, raise)) {
            const float nsh = raise ? cm : sh;
            if (c0 > 0) {                                      // rescale what this row has already written
              const float f = (sh == -INFINITY) ? 0.f : ex2f((sh - nsh) * c);     // 1 for rows whose shift stays
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              for (int pc = 0; pc < c0; pc += 32) {
                float q[16];
                tmem_ld16(tmem + lane_off + (pc >> 1), q);
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const uint32_t u = __float_as_uint(q[j]);
                  const float2 x = __bfloat1622float2(*(const __nv_bfloat162*)&u);
                  w[j] = pack_bf16(x.x * f, x.y * f);
                }
                tmem_st16(tmem + lane_off + (pc >> 1), w);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) s4[k] *= f;
            }
            sh = nsh;
          }
          const float shc = (sh == -INFINITY) ? 0.f : sh * c;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2f(fmaf(s[j], c, -shc)), p1 = ex2f(fmaf(s[j + 1], c, -shc));
            s4[(j >> 1) & 3] += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
          tmem_st16(tmem + lane_off + (c0 >> 1), pk);
        };
        tmem_ld32_nowait(tmem + lane_off, bufa);
        for (int c0 = 0; c0 < npad; c0 += 64) {
          online_chunk(bufa, bufb, c0);
          if (c0 + 32 < npad) online_chunk(bufb, bufa, c0 + 32);
        }
        m = sh;
      } else {
      // Both passes keep the next 32-column TMEM load in flight while the current one is processed (two register
      // buffers) and use 4 independent max / sum chains.
      // ---- pass 1: row max
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      auto max_chunk = [&](float (&s)[32], float (&nxt)[32], int c0) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 < npad) tmem_ld32_nowait(tmem + lane_off + c0 + 32, nxt);
        if (c0 + 32 <= kmax) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], c0 + j < kmax ? s[j] : -INFINITY);
        }
      };
      tmem_ld32_nowait(tmem + lane_off, bufa);
      for (int c0 = 0; c0 < npad; c0 += 64) {
        max_chunk(bufa, bufb, c0);
        if (c0 + 32 < npad) max_chunk(bufb, bufa, c0 + 32);
      }
      m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      if (warp == 0) TRACE(1, 32);
      const float mc = (m == -INFINITY) ? 0.f : m * c;        // dead rows (qi >= L) only
      // ---- pass 2: p = exp2(s c - m c), row sum, packed bf16 P over the consumed S columns
      auto exp_chunk = [&](float (&s)[32], float (&nxt)[32], int c0) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // this chunk's P lands on columns [c0/2, c0/2+16), all below c0+32: it never touches S that is still unread
        if (c0 + 32 < npad) tmem_ld32_nowait(tmem + lane_off + c0 + 32, nxt);
        uint32_t pk[16];
        if (c0 + 32 <= kmax) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2f(fmaf(s[j], c, -mc)), p1 = ex2f(fmaf(s[j + 1], c, -mc));
            s4[(j >> 1) & 3] += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = c0 + j < kmax ? ex2f(fmaf(s[j], c, -mc)) : 0.f;
            const float p1 = c0 + j + 1 < kmax ? ex2f(fmaf(s[j + 1], c, -mc)) : 0.f;
            s4[(j >> 1) & 3] += p0 + p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
        }
        tmem_st16(tmem + lane_off + (c0 >> 1), pk);
      };
      tmem_ld32_nowait(tmem + lane_off, bufa);
      for (int c0 = 0; c0 < npad; c0 += 64) {
        exp_chunk(bufa, bufb, c0);
        if (c0 + 32 < npad) exp_chunk(bufb, bufa, c0 + 32);
      }
      }
      sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    if (warp == 0) TRACE(1, 33);
    mbar_wait(bar_o, par);
    tcgen05_fence_after();
    if (warp == 0) TRACE(1, 34);
    float o[64];
    if (warp_live) {
      tmem_ld32_nowait(tmem + lane_off + F_TM_O, o);
      tmem_ld32_nowait(tmem + lane_off + F_TM_O + 32, o + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_free);
    if (warp == 0) TRACE(1, 35);
    if (warp_live) {
      // O leaves through the TMA (32 x 64 box per warp over a 3-D map {H*64, L, B}: rows past L are clipped): per-lane
      // 128-byte row stores cost 32 LSU wavefronts per instruction, ~2.4 k cycles per item (SC_ATT_TRACE)
      const float inv = 1.0f / sum;
      const uint32_t ostage = sbase + F_SM_O + warp * 4096;
      bulk_wait_read<0>();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128u(ostage + lane * 128 + ((j ^ (lane & 7)) << 4), pack_bf16(o[8 * j] * inv, o[8 * j + 1] * inv),
                pack_bf16(o[8 * j + 2] * inv, o[8 * j + 3] * inv), pack_bf16(o[8 * j + 4] * inv, o[8 * j + 5] * inv),
                pack_bf16(o[8 * j + 6] * inv, o[8 * j + 7] * inv));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (!pad) {
        asm volatile(
            "{\n\t.reg .pred e;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "@e cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n\t}"
            ::"l"((uint64_t)&tmO), "r"(ostage), "r"(h * HD), "r"(qt * TILE + warp * 32), "r"(b) : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred e;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "@e cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n\t}"
            ::"l"((uint64_t)&tmO), "r"(ostage), "r"(0), "r"(h), "r"(qt * TILE + warp * 32), "r"(b) : "memory");
      }
      bulk_commit();
      if (qi < L) a.lse[((long)b * a.H + h) * L + qi] = m * a.scale + logf(sum);
    }
    if (warp == 0) TRACE(1, 36);
    }   // item loop
    if constexpr (F_NSOFT == 4) bulk_wait_all();           // staging smem must outlive the last TMA stores
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == F_NSOFT) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(F_TM_COLS) : "memory");
  }
}


}  // namespace

extern void sc_count_kernel(int kind, int n);

// Tensor maps of a head-padded operand (head dim 32 / 48): loads {hd, H, rows} with a 64 x 1 x box_rows box (zero fill past the
// head), stores {hd, H, L, B} with a box_cols x 1 x 32 x 1 box (columns past the head and rows past the sample clipped).
static int pad_load_map(const void* p, int hd, int H, long rows, long rs, uint32_t box_rows, CUtensorMap* out) {
  const uint64_t dims[3] = {(uint64_t)hd, (uint64_t)H, (uint64_t)rows}, str[2] = {(uint64_t)hd, (uint64_t)rs};
  const uint32_t box[3] = {64, 1, box_rows};
  return sc_get_tensor_map_nd(p, 3, dims, str, box, 128, out);
}
static int pad_store_map(const void* p, int hd, int H, int L, int B, long rs, long bs, uint32_t box_cols, int swizzle, CUtensorMap* out) {
  const uint64_t dims[4] = {(uint64_t)hd, (uint64_t)H, (uint64_t)L, (uint64_t)B}, str[3] = {(uint64_t)hd, (uint64_t)rs, (uint64_t)bs};
  const uint32_t box[4] = {box_cols, 1, 32, 1};
  return sc_get_tensor_map_nd(p, 4, dims, str, box, swizzle, out);
}

bool sc_attn_tc_supported(const sc_attn_desc* a) {
  auto packed = [&](const void* p, long bs, long rs, int L) {
    return ((uintptr_t)p & 15) == 0 && rs % 8 == 0 && bs == (long)L * rs;
  };
  static const bool pad_ok = getenv("SC_ATT_TC_PAD") ? atoi(getenv("SC_ATT_TC_PAD")) != 0 : true;   // A/B switch: head dims < 64 on tcgen05
  const bool hd_ok = a->hd == 64 || (pad_ok && (a->hd == 48 || a->hd == 32) && a->q_rs == a->k_rs && a->q_rs == a->v_rs);
  return a->dtype == SC_BF16 && hd_ok && a->Lq == a->Lk && a->Lq >= 16 && a->Lq <= 256 && a->B <= 65535 &&
         packed(a->q, a->q_bs, a->q_rs, a->Lq) && packed(a->k, a->k_bs, a->k_rs, a->Lk) &&
         packed(a->v, a->v_bs, a->v_rs, a->Lk) && packed(a->o, a->o_bs, a->o_rs, a->Lq);
}

// delta[b,h,i] = sum_d dO[b,i,h,d] O[b,i,h,d] (head dim 64) as a stand-alone pass; used by sc_gemm for dot_out when its kernel
// cannot fuse the dots
int sc_attn_delta(const void* o, const void* d_o, long bs, long rs, int B, int H, int L, float* delta, cudaStream_t st) {
  const long n = (long)B * L * H;
  attn_delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const bf16*)o, (const bf16*)d_o, bs, rs, B, H, L, delta);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_attention_bwd_tc(const sc_attn_bwd_desc* g, float* delta, cudaStream_t st) {
  const sc_attn_desc* a = &g->fwd;
  const int L = a->Lq;
  const long rows = (long)a->B * L;
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  CUtensorMap gq, gk, gv;
  if (a->hd == HD) {
    // 2-D maps over the [B*L, H*64] column slices; box = 64 columns x 128 rows, zero fill past the last row
    if ((rc = sc_get_tensor_map(a->q, (uint64_t)a->H * HD, rows, a->q_rs, 64, TILE, &tq))) return rc;
    if ((rc = sc_get_tensor_map(a->k, (uint64_t)a->H * HD, rows, a->k_rs, 64, TILE, &tk))) return rc;
    if ((rc = sc_get_tensor_map(a->v, (uint64_t)a->H * HD, rows, a->v_rs, 64, TILE, &tv))) return rc;
    if ((rc = sc_get_tensor_map(g->d_o, (uint64_t)a->H * HD, rows, a->o_rs, 64, TILE, &tdo))) return rc;
    // gradient outputs: 3-D {H*64, L, B}, 32 x 32 boxes, rows past L clipped
    if ((rc = sc_get_tensor_map_3d(g->d_q, (uint64_t)a->H * HD, L, a->B, a->q_rs, a->q_bs, 32, 32, 64, &gq))) return rc;
    if ((rc = sc_get_tensor_map_3d(g->d_k, (uint64_t)a->H * HD, L, a->B, a->k_rs, a->k_bs, 32, 32, 64, &gk))) return rc;
    if ((rc = sc_get_tensor_map_3d(g->d_v, (uint64_t)a->H * HD, L, a->B, a->v_rs, a->v_bs, 32, 32, 64, &gv))) return rc;
  } else {
    if ((rc = pad_load_map(a->q, a->hd, a->H, rows, a->q_rs, TILE, &tq))) return rc;
    if ((rc = pad_load_map(a->k, a->hd, a->H, rows, a->k_rs, TILE, &tk))) return rc;
    if ((rc = pad_load_map(a->v, a->hd, a->H, rows, a->v_rs, TILE, &tv))) return rc;
    if ((rc = pad_load_map(g->d_o, a->hd, a->H, rows, a->o_rs, TILE, &tdo))) return rc;
    if ((rc = pad_store_map(g->d_q, a->hd, a->H, L, a->B, a->q_rs, a->q_bs, 32, 64, &gq))) return rc;
    if ((rc = pad_store_map(g->d_k, a->hd, a->H, L, a->B, a->k_rs, a->k_bs, 32, 64, &gk))) return rc;
    if ((rc = pad_store_map(g->d_v, a->hd, a->H, L, a->B, a->v_rs, a->v_bs, 32, 64, &gv))) return rc;
  }
  sc_count_kernel(SC_K_ATTN_BWD_TC, g->delta_ready ? 1 : 2);
  if (!g->delta_ready) {         // (otherwise the out_proj dgrad that produced dO has already written it: sc_gemm dot_out)
    const long n = rows * a->H;
    attn_delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const bf16*)a->o, (const bf16*)g->d_o, a->o_bs, a->o_rs, a->B,
                                                                   a->H, L, delta, a->hd);
  }
  const long items = (long)a->H * a->B;
  dim3 grid((unsigned)(items < sc_num_sms() ? items : sc_num_sms()));       // persistent: one CTA per SM
  if (a->causal) {
    static sc_device_once once;
    if (once.first()) { SC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL)); once.done(); }
    SC_CUDA(sc_launch_pdl(attn_bwd_tc_kernel<true>, grid, dim3(TC_THREADS), SMEM_TOTAL, st, tq, tk, tv, tdo, gq, gk, gv, *g, delta));
  } else {
    static sc_device_once once;
    if (once.first()) { SC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL)); once.done(); }
    SC_CUDA(sc_launch_pdl(attn_bwd_tc_kernel<false>, grid, dim3(TC_THREADS), SMEM_TOTAL, st, tq, tk, tv, tdo, gq, gk, gv, *g, delta));
  }
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_attention_fwd_tc(const sc_attn_desc* a, cudaStream_t st) {
  const int L = a->Lq, ntile = (L + TILE - 1) / TILE;
  const long rows = (long)a->B * L;
  CUtensorMap tq, tk, tv;
  int rc;
  CUtensorMap to;
  if (a->hd == HD) {
    if ((rc = sc_get_tensor_map(a->q, (uint64_t)a->H * HD, rows, a->q_rs, 64, TILE, &tq))) return rc;
    if ((rc = sc_get_tensor_map(a->k, (uint64_t)a->H * HD, rows, a->k_rs, 64, ntile * TILE, &tk))) return rc;
    if ((rc = sc_get_tensor_map(a->v, (uint64_t)a->H * HD, rows, a->v_rs, 64, ntile * TILE, &tv))) return rc;
    if ((rc = sc_get_tensor_map_3d(a->o, (uint64_t)a->H * HD, L, a->B, a->o_rs, a->o_bs, 64, 32, 128, &to))) return rc;
  } else {
    if ((rc = pad_load_map(a->q, a->hd, a->H, rows, a->q_rs, TILE, &tq))) return rc;
    if ((rc = pad_load_map(a->k, a->hd, a->H, rows, a->k_rs, ntile * TILE, &tk))) return rc;
    if ((rc = pad_load_map(a->v, a->hd, a->H, rows, a->v_rs, ntile * TILE, &tv))) return rc;
    if ((rc = pad_store_map(a->o, a->hd, a->H, L, a->B, a->o_rs, a->o_bs, 64, 128, &to))) return rc;
  }
  sc_count_kernel(SC_K_ATTN_FWD_TC, 1);
  const long total = (long)ntile * a->H * a->B;
  static const int online = getenv("SC_ATT_FWD_TWO_PASS") ? 0 : 1; // A/B switch: exact row maximum first (reads S twice)
  const long slots = 2L * sc_num_sms();
  dim3 grid((unsigned)(total < slots ? total : slots));
  if (a->causal) {
    static sc_device_once once;
    if (once.first()) { SC_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_TOTAL)); once.done(); }
    SC_CUDA(sc_launch_pdl(attn_fwd_tc_kernel<true>, grid, dim3(F_THREADS), F_SMEM_TOTAL, st, tq, tk, tv, to, *a, online));
  } else {
    static sc_device_once once;
    if (once.first()) { SC_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_TOTAL)); once.done(); }
    SC_CUDA(sc_launch_pdl(attn_fwd_tc_kernel<false>, grid, dim3(F_THREADS), F_SMEM_TOTAL, st, tq, tk, tv, to, *a, online));
  }
  SC_LAUNCH_CHECK();
  return SC_OK;
}

#ifdef SC_ATT_TRACE
extern "C" int sc_debug_attn_trace(long long* out, int* n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 2 * 512);
  cudaMemcpyFromSymbol(n, g_trace_n, sizeof(int) * 2);
  int z[2] = {0, 0};
  cudaMemcpyToSymbol(g_trace_n, z, sizeof(z));
  return 0;
}
#endif
