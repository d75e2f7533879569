// LayerNorm forward / backward (one warp per row, the row lives in registers).
// Replaces modules/module_clip_util.py:126-132 (LayerNorm, fp32 statistics) and the nn.LayerNorm
// instances in modules/module_seg_vit.py:256,264,267,272 and modules/module_mae.py:146,149,239.
// HBM-bound: read x once, write y once (fwd); read dy, x once, write dx once (bwd).
#include <stdlib.h>

#include "common.cuh"

namespace {

template <typename T> SC_DEVINL float4 load4(const T* p, long i4);
template <> SC_DEVINL float4 load4<float>(const float* p, long i4) { return *(const float4*)(p + i4 * 4); }
template <> SC_DEVINL float4 load4<bf16>(const bf16* p, long i4) {
  uint2 u = *(const uint2*)(p + i4 * 4);
  __nv_bfloat162 a = *(__nv_bfloat162*)&u.x, b = *(__nv_bfloat162*)&u.y;
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> SC_DEVINL void store4(T* p, long i4, float4 v);
template <> SC_DEVINL void store4<float>(float* p, long i4, float4 v) { *(float4*)(p + i4 * 4) = v; }
template <> SC_DEVINL void store4<bf16>(bf16* p, long i4, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *(uint32_t*)&a;
  u.y = *(uint32_t*)&b;
  *(uint2*)(p + i4 * 4) = u;
}

SC_DEVINL long remap_row(long r, int in_group, int out_group, int out_off) {
  if (in_group <= 0) return r;
  return (r / in_group) * (long)out_group + out_off + (r % in_group);
}

template <typename TX, typename TY, int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(sc_ln_desc d) {
  pdl_launch_dependents();      // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= d.rows) return;
  const int D4 = d.D >> 2;
  const TX* x = (const TX*)d.x + row * d.D;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    int i4 = lane + 32 * j;
    v[j] = (i4 < D4) ? load4<TX>(x, i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  const float mean = warp_sum(s) / d.D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    int i4 = lane + 32 * j;
    if (i4 < D4) {
      float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
      q += a * a + b * b + c * c + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d.D + d.eps);
  if (lane == 0) {
    if (d.mean) d.mean[row] = mean;
    if (d.rstd) d.rstd[row] = rstd;
  }
  TY* y = (TY*)d.y + remap_row(row, d.in_group, d.out_group, d.out_off) * d.D;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    int i4 = lane + 32 * j;
    if (i4 < D4) {
      float4 g = *(const float4*)(d.gamma + i4 * 4), b = *(const float4*)(d.beta + i4 * 4);
      float4 o;
      o.x = (v[j].x - mean) * rstd * g.x + b.x;
      o.y = (v[j].y - mean) * rstd * g.y + b.y;
      o.z = (v[j].z - mean) * rstd * g.z + b.z;
      o.w = (v[j].w - mean) * rstd * g.w + b.w;
      store4<TY>(y, i4, o);
    }
  }
}

// Backward: one warp per row (grid-stride).  Each warp keeps its dgamma / dbeta partial sums in a PRIVATE slice of
// shared memory (plain vector load-add-store, no atomics, no conflicts) instead of 2*NV*4 registers per thread: the
// kernel stays under ~80 registers so three CTAs (24 warps) are resident per SM and enough reads are in flight for HBM.
template <typename TDY, typename TX, typename TDX, int NV>
__global__ void __launch_bounds__(256, 3) ln_bwd_kernel(sc_ln_bwd_desc d) {
  extern __shared__ __align__(16) float sacc[];          // [8 warps][3][D]: dgamma, dbeta, colsum(dx)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D4 = d.D >> 2;
  const bool want_param = d.dgamma != nullptr;
  const bool want_cs = d.dx_colsum != nullptr;
  const int nacc = want_cs ? 3 : 2;                      // arrays per warp (the colsum slice only when requested)
  float* mine = sacc + (size_t)warp * nacc * d.D;
  if (want_param || want_cs) {
    for (int i = lane; i < nacc * d.D; i += 32) mine[i] = 0.f;
    __syncwarp();
  }
  const bool acc = d.dx && d.accumulate_dx;
  for (long row = (long)blockIdx.x * 8 + warp; row < d.rows; row += (long)gridDim.x * 8) {
    const TX* x = (const TX*)d.x + row * d.D;
    const TDY* dy = (const TDY*)d.dy + remap_row(row, d.in_group, d.out_group, d.out_off) * d.D;
    float4 xv[NV], dv[NV], old[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {      // every global read of the row is issued up front
      const int i4 = lane + 32 * j;
      const bool in = i4 < D4;
      xv[j] = in ? load4<TX>(x, i4) : make_float4(0.f, 0.f, 0.f, 0.f);
      dv[j] = in ? load4<TDY>(dy, i4) : make_float4(0.f, 0.f, 0.f, 0.f);
      old[j] = (acc && in) ? load4<TDX>((const TDX*)d.dx + row * d.D, i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mean = d.mean[row], rstd = d.rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i4 = lane + 32 * j;
      if (i4 < D4) {
        const float4 gm = __ldg((const float4*)(d.gamma + i4 * 4));
        xv[j] = make_float4((xv[j].x - mean) * rstd, (xv[j].y - mean) * rstd, (xv[j].z - mean) * rstd, (xv[j].w - mean) * rstd);
        const float gx = dv[j].x * gm.x, gy = dv[j].y * gm.y, gz = dv[j].z * gm.z, gw = dv[j].w * gm.w;
        s1 += gx + gy + gz + gw;
        s2 += gx * xv[j].x + gy * xv[j].y + gz * xv[j].z + gw * xv[j].w;
        if (want_param) {
          float4* sg = (float4*)(mine + i4 * 4);
          float4* sb = (float4*)(mine + d.D + i4 * 4);
          float4 a = *sg, b = *sb;
          a.x += dv[j].x * xv[j].x; a.y += dv[j].y * xv[j].y; a.z += dv[j].z * xv[j].z; a.w += dv[j].w * xv[j].w;
          b.x += dv[j].x; b.y += dv[j].y; b.z += dv[j].z; b.w += dv[j].w;
          *sg = a;
          *sb = b;
        }
      }
    }
    const float c1 = warp_sum(s1) / d.D, c2 = warp_sum(s2) / d.D;
    if (d.dx) {
      TDX* dx = (TDX*)d.dx + row * d.D;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int i4 = lane + 32 * j;
        if (i4 < D4) {
          const float4 gm = __ldg((const float4*)(d.gamma + i4 * 4));
          float4 o;
          o.x = rstd * (dv[j].x * gm.x - c1 - xv[j].x * c2) + old[j].x;
          o.y = rstd * (dv[j].y * gm.y - c1 - xv[j].y * c2) + old[j].y;
          o.z = rstd * (dv[j].z * gm.z - c1 - xv[j].z * c2) + old[j].z;
          o.w = rstd * (dv[j].w * gm.w - c1 - xv[j].w * c2) + old[j].w;
          store4<TDX>(dx, i4, o);
          if (d.dx_copy_bf16) store4<bf16>((bf16*)d.dx_copy_bf16 + row * d.D, i4, o);
          if (want_cs) {
            float4* sc = (float4*)(mine + 2 * d.D + i4 * 4);
            float4 a = *sc;
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            *sc = a;
          }
        }
      }
    }
  }
  if (!want_param && !want_cs) return;
  __syncthreads();
  for (int i = threadIdx.x; i < nacc * d.D; i += 256) {
    float* dst = i < d.D ? d.dgamma : (i < 2 * d.D ? d.dbeta : d.dx_colsum);
    if (dst == nullptr) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sacc[(size_t)w * nacc * d.D + i];
    atomicAdd(dst + (i % d.D), s);
  }
}


// ---- register-accumulator backward (D % 8 == 0): the production kernel ---------------------------------------------
// One warp per row, 8-element vectors (16-byte loads of bf16 data).  dgamma / dbeta / colsum(dx) partial sums live in
// REGISTERS (8 * NV8 each) for the whole grid-stride loop and are reduced across the CTA's warps once at the end.  The
// first version kept them in per-warp shared-memory slices (2-3 vector load-add-store per 4 elements): with the bf16
// gradient stream the kernel moves 10 B per element instead of 16 and that smem traffic, not HBM, became the limit
// (3.0 TB/s).  128-thread CTAs, 3 per SM.
template <typename T> SC_DEVINL void ld8(const T* p, float (&v)[8]);
template <> SC_DEVINL void ld8<float>(const float* p, float (&v)[8]) {
  const float4 a = *(const float4*)p, b = *(const float4*)(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> SC_DEVINL void ld8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = *(const uint4*)p;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*(const __nv_bfloat162*)&w[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <typename T> SC_DEVINL void st8(T* p, const float (&v)[8]);
template <> SC_DEVINL void st8<float>(float* p, const float (&v)[8]) {
  *(float4*)p = make_float4(v[0], v[1], v[2], v[3]);
  *(float4*)(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> SC_DEVINL void st8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), e = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *(uint32_t*)&a; u.y = *(uint32_t*)&b; u.z = *(uint32_t*)&c; u.w = *(uint32_t*)&e;
  *(uint4*)p = u;
}

constexpr int LNB_WARPS = 4;

template <typename TDY, typename TX, typename TDX, int NV8, bool CS>
__global__ void __launch_bounds__(LNB_WARPS * 32, NV8 >= 4 ? 2 : 3) ln_bwd8_kernel(sc_ln_bwd_desc d) {
  __shared__ float red[LNB_WARPS][1024];
  pdl_launch_dependents();      // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D8 = d.D >> 3;
  const bool want_param = d.dgamma != nullptr;
  const bool acc = d.dx && d.accumulate_dx;
  float dg[NV8][8], db[NV8][8], cs[CS ? NV8 : 1][8];
#pragma unroll
  for (int j = 0; j < NV8; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dg[j][k] = 0.f;
      db[j][k] = 0.f;
      if (CS) cs[j][k] = 0.f;
    }
  for (long row = (long)blockIdx.x * LNB_WARPS + warp; row < d.rows; row += (long)gridDim.x * LNB_WARPS) {
    const TX* x = (const TX*)d.x + row * d.D;
    const TDY* dy = (const TDY*)d.dy + remap_row(row, d.in_group, d.out_group, d.out_off) * d.D;
    float xv[NV8][8], dv[NV8][8], old[NV8][8];
#pragma unroll
    for (int j = 0; j < NV8; ++j) {      // every global read of the row is issued up front
      const int c8 = lane + 32 * j;
      if (c8 < D8) {
        ld8<TX>(x + c8 * 8, xv[j]);
        ld8<TDY>(dy + c8 * 8, dv[j]);
        if (acc) ld8<TDX>((const TDX*)d.dx + row * d.D + c8 * 8, old[j]);
      }
    }
    const float mean = d.mean[row], rstd = d.rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV8; ++j) {
      const int c8 = lane + 32 * j;
      if (c8 < D8) {
        float gm[8];
        ld8<float>(d.gamma + c8 * 8, gm);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (xv[j][k] - mean) * rstd;
          const float g = dv[j][k] * gm[k];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          dg[j][k] = fmaf(dv[j][k], xh, dg[j][k]);
          db[j][k] += dv[j][k];
          xv[j][k] = xh;
          dv[j][k] = g;
        }
      }
    }
    const float c1 = warp_sum(s1) / d.D, c2 = warp_sum(s2) / d.D;
    if (d.dx) {
      TDX* dx = (TDX*)d.dx + row * d.D;
#pragma unroll
      for (int j = 0; j < NV8; ++j) {
        const int c8 = lane + 32 * j;
        if (c8 < D8) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o[k] = rstd * (dv[j][k] - c1 - xv[j][k] * c2);
            if (acc) o[k] += old[j][k];
            if (CS) cs[j][k] += o[k];
          }
          st8<TDX>(dx + c8 * 8, o);
          if (d.dx_copy_bf16) st8<bf16>((bf16*)d.dx_copy_bf16 + row * d.D + c8 * 8, o);
        }
      }
    }
  }
  if (!want_param && !CS) return;
  // cross-warp reduction, one array at a time: registers -> smem slice per warp -> one atomic per column and CTA
#pragma unroll
  for (int which = 0; which < (CS ? 3 : 2); ++which) {
    float* dst = which == 0 ? d.dgamma : (which == 1 ? d.dbeta : d.dx_colsum);
    if (dst == nullptr) continue;          // uniform
#pragma unroll
    for (int j = 0; j < NV8; ++j) {
      const int c8 = lane + 32 * j;
      if (c8 < D8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red[warp][c8 * 8 + k] = which == 0 ? dg[j][k] : (which == 1 ? db[j][k] : cs[CS ? j : 0][k]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d.D; i += LNB_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < LNB_WARPS; ++w) s += red[w][i];
      atomicAdd(dst + i, s);
    }
    __syncthreads();
  }
}

template <typename TDY, typename TX, typename TDX>
int launch_bwd8(const sc_ln_bwd_desc& d, cudaStream_t st) {
  const int nv = ceil_div(d.D, 256);
  long g = ceil_div(d.rows, LNB_WARPS);
  const long per_sm = nv >= 4 ? 2 : 3;                                         // resident CTAs per SM (register budget)
  const int grid = (int)(g < per_sm * sc_num_sms() ? g : per_sm * sc_num_sms());
#define SC_LN8_CASE(NV_)                                                                     \
  if (d.dx_colsum) sc_launch_pdl(ln_bwd8_kernel<TDY, TX, TDX, NV_, true>, dim3(grid), dim3(LNB_WARPS * 32), 0, st, d); \
  else sc_launch_pdl(ln_bwd8_kernel<TDY, TX, TDX, NV_, false>, dim3(grid), dim3(LNB_WARPS * 32), 0, st, d);
  switch (nv) {
    case 1: SC_LN8_CASE(1); break;
    case 2: SC_LN8_CASE(2); break;
    case 3: SC_LN8_CASE(3); break;
    default: SC_LN8_CASE(4); break;
  }
#undef SC_LN8_CASE
  SC_LAUNCH_CHECK();
  return SC_OK;
}

// ---- bulk-copy-prefetched backward (the production kernel for D % 8 == 0) -------------------------------------------
// ln_bwd8_kernel above keeps a whole row per warp in registers: with 12 warps per SM only 12 rows (~90 KB) are in flight
// per SM and the kernel is bound by HBM LATENCY (4.0 TB/s measured, 10 B per element).  Here every warp owns a ring of
// LNT_STAGES row slots in shared memory that is filled by cp.async.bulk (the TMA's 1-D copy: x row, dy row, old dx row,
// one mbarrier per slot) up to three rows ahead: 36 rows (216 KB) in flight per SM at no register cost.  The warp makes
// two passes over its slot (statistics, then outputs), so only the three accumulator sets live in registers.
constexpr int LNT_WARPS = 6;      // 2 CTAs x 6 warps x 3 slots of 6 KB (D = 768) fill the SM's shared memory
constexpr int LNT_STAGES = 3;

SC_DEVINL uint32_t lnt_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SC_DEVINL void lnt_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
SC_DEVINL void lnt_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
template <typename T> SC_DEVINL void lds8(uint32_t a, float (&v)[8]);
template <> SC_DEVINL void lds8<float>(uint32_t a, float (&v)[8]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a + 16));
}
template <> SC_DEVINL void lds8<bf16>(uint32_t a, float (&v)[8]) {
  uint32_t w[4];
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*(const __nv_bfloat162*)&w[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

template <typename TDY, typename TX, typename TDX, int NV8, bool CS>
__global__ void __launch_bounds__(LNT_WARPS * 32, 2) ln_bwd_tma_kernel(sc_ln_bwd_desc d) {
  extern __shared__ __align__(128) uint8_t lsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = d.D, D8 = D >> 3;
  const bool want_param = d.dgamma != nullptr;
  const bool acc = d.dx && d.accumulate_dx;
  const uint32_t xb = D * sizeof(TX), yb = D * sizeof(TDY), ob = acc ? D * (uint32_t)sizeof(TDX) : 0u;
  const uint32_t slot_bytes = xb + yb + ob;                  // multiples of 16 (D % 8 == 0)
  // layout: gamma [D] | per warp: LNT_STAGES slots | mbarriers; the cross-warp reduction at the end reuses the slots
  float* gam = (float*)lsm;
  uint8_t* ring = lsm + D * 4 + (size_t)warp * LNT_STAGES * slot_bytes;
  uint64_t* bars = (uint64_t*)(lsm + D * 4 + (size_t)LNT_WARPS * LNT_STAGES * slot_bytes) + warp * LNT_STAGES;
  pdl_launch_dependents();      // programmatic dependent launch (common.cuh): barrier init under the previous kernel's tail
  if (lane == 0) {
    for (int s = 0; s < LNT_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lnt_smem(&bars[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();
  for (int i = threadIdx.x; i < D; i += LNT_WARPS * 32) gam[i] = d.gamma[i];
  __syncthreads();
  const long stride = (long)gridDim.x * LNT_WARPS;
  const long row0 = (long)blockIdx.x * LNT_WARPS + warp;
  auto issue = [&](long row, int s) {                        // lane 0 only
    const uint32_t bar = lnt_smem(&bars[s]);
    const uint32_t dst = lnt_smem(ring + (size_t)s * slot_bytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slot_bytes) : "memory");
    lnt_bulk(dst, (const TX*)d.x + row * D, xb, bar);
    lnt_bulk(dst + xb, (const TDY*)d.dy + remap_row(row, d.in_group, d.out_group, d.out_off) * D, yb, bar);
    if (acc) lnt_bulk(dst + xb + yb, (const TDX*)d.dx + row * D, ob, bar);
  };
  if (lane == 0) {
    for (int s = 0; s < LNT_STAGES; ++s)
      if (row0 + s * stride < d.rows) issue(row0 + s * stride, s);
  }
  float dg[NV8][8], db[NV8][8], cs[CS ? NV8 : 1][8];
#pragma unroll
  for (int j = 0; j < NV8; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dg[j][k] = 0.f;
      db[j][k] = 0.f;
      if (CS) cs[j][k] = 0.f;
    }
  const uint32_t gbase = lnt_smem(gam);
  int s = 0;
  uint32_t phase = 0;
  for (long row = row0; row < d.rows; row += stride) {
    const float mean = d.mean[row], rstd = d.rstd[row];      // issued before the wait: overlaps the copy's flight
    const uint32_t slot = lnt_smem(ring + (size_t)s * slot_bytes);
    lnt_wait(lnt_smem(&bars[s]), phase);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV8; ++j) {
      const int c8 = lane + 32 * j;
      if (c8 < D8) {
        float xv[8], dv[8], gm[8];
        lds8<TX>(slot + c8 * 8 * (uint32_t)sizeof(TX), xv);
        lds8<TDY>(slot + xb + c8 * 8 * (uint32_t)sizeof(TDY), dv);
        lds8<float>(gbase + c8 * 32, gm);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (xv[k] - mean) * rstd;
          const float g = dv[k] * gm[k];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          dg[j][k] = fmaf(dv[k], xh, dg[j][k]);
          db[j][k] += dv[k];
        }
      }
    }
    const float c1 = warp_sum(s1) / D, c2 = warp_sum(s2) / D;
    if (d.dx) {
      TDX* dx = (TDX*)d.dx + row * D;
#pragma unroll
      for (int j = 0; j < NV8; ++j) {
        const int c8 = lane + 32 * j;
        if (c8 < D8) {
          float xv[8], dv[8], gm[8], o[8];
          lds8<TX>(slot + c8 * 8 * (uint32_t)sizeof(TX), xv);
          lds8<TDY>(slot + xb + c8 * 8 * (uint32_t)sizeof(TDY), dv);
          lds8<float>(gbase + c8 * 32, gm);
          if (acc) lds8<TDX>(slot + xb + yb + c8 * 8 * (uint32_t)sizeof(TDX), o);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float xh = (xv[k] - mean) * rstd;
            const float v = rstd * (dv[k] * gm[k] - c1 - xh * c2);
            o[k] = acc ? o[k] + v : v;
            if (CS) cs[j][k] += o[k];
          }
          st8<TDX>(dx + c8 * 8, o);
          if (d.dx_copy_bf16) st8<bf16>((bf16*)d.dx_copy_bf16 + row * D + c8 * 8, o);
        }
      }
    }
    __syncwarp();                                            // every lane has read the slot
    if (lane == 0 && row + LNT_STAGES * stride < d.rows) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(row + LNT_STAGES * stride, s);
    }
    if (++s == LNT_STAGES) { s = 0; phase ^= 1; }
  }
  if (!want_param && !CS) return;
  __syncthreads();                                           // all rings idle: reuse them as the reduction scratch
  float* red = (float*)(lsm + D * 4);                        // [LNT_WARPS][D]
#pragma unroll
  for (int which = 0; which < (CS ? 3 : 2); ++which) {
    float* dst = which == 0 ? d.dgamma : (which == 1 ? d.dbeta : d.dx_colsum);
    if (dst == nullptr) continue;          // uniform
#pragma unroll
    for (int j = 0; j < NV8; ++j) {
      const int c8 = lane + 32 * j;
      if (c8 < D8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red[warp * D + c8 * 8 + k] = which == 0 ? dg[j][k] : (which == 1 ? db[j][k] : cs[CS ? j : 0][k]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += LNT_WARPS * 32) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < LNT_WARPS; ++w) sum += red[w * D + i];
      atomicAdd(dst + i, sum);
    }
    __syncthreads();
  }
}

template <typename TDY, typename TX, typename TDX>
int launch_bwd_tma(const sc_ln_bwd_desc& d, cudaStream_t st) {
  const int nv = ceil_div(d.D, 256);
  const bool acc = d.dx && d.accumulate_dx;
  const size_t slot = (size_t)d.D * (sizeof(TX) + sizeof(TDY) + (acc ? sizeof(TDX) : 0));
  size_t smem = (size_t)d.D * 4 + LNT_WARPS * LNT_STAGES * slot + LNT_WARPS * LNT_STAGES * 8;
  const size_t red = (size_t)d.D * 4 + (size_t)LNT_WARPS * d.D * 4;
  if (smem < red) smem = red;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) return SC_ERR_UNSUPPORTED;
  long g = ceil_div(d.rows, LNT_WARPS);
  const int grid = (int)(g < (long)per_sm * sc_num_sms() ? g : (long)per_sm * sc_num_sms());
#define SC_LNT_CASE(NV_, CS_)                                                                                            \
  {                                                                                                                      \
    static sc_device_once once;                                                                                          \
    if (once.first()) { cudaFuncSetAttribute(ln_bwd_tma_kernel<TDY, TX, TDX, NV_, CS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); once.done(); } \
    sc_launch_pdl(ln_bwd_tma_kernel<TDY, TX, TDX, NV_, CS_>, dim3(grid), dim3(LNT_WARPS * 32), smem, st, d);              \
  }
#define SC_LNT_NV(NV_) if (d.dx_colsum) SC_LNT_CASE(NV_, true) else SC_LNT_CASE(NV_, false)
  switch (nv) {
    case 1: SC_LNT_NV(1); break;
    case 2: SC_LNT_NV(2); break;
    case 3: SC_LNT_NV(3); break;
    default: SC_LNT_NV(4); break;
  }
#undef SC_LNT_NV
#undef SC_LNT_CASE
  SC_LAUNCH_CHECK();
  return SC_OK;
}

template <typename TX, typename TY>
int launch_fwd(const sc_ln_desc& d, cudaStream_t st) {
  const int nv = ceil_div(d.D, 128);
  const int grid = ceil_div(d.rows, 8);
#define SC_LN_CASE(NV_) sc_launch_pdl(ln_fwd_kernel<TX, TY, NV_>, dim3(grid), dim3(256), 0, st, d)
  switch (nv) {
    case 1: SC_LN_CASE(1); break;
    case 2: SC_LN_CASE(2); break;
    case 3: SC_LN_CASE(3); break;
    case 4: SC_LN_CASE(4); break;
    case 5: case 6: SC_LN_CASE(6); break;
    case 7: case 8: SC_LN_CASE(8); break;
    default: sc_set_error("sc_layernorm_fwd: D=%d > 1024 unsupported", d.D); return SC_ERR_UNSUPPORTED;
  }
#undef SC_LN_CASE
  SC_LAUNCH_CHECK();
  return SC_OK;
}

template <typename TDY, typename TX, typename TDX>
int launch_bwd(const sc_ln_bwd_desc& d, cudaStream_t st) {
  const int nv = ceil_div(d.D, 128);
  long g = ceil_div(d.rows, 8);
  const int grid = (int)(g < 3L * sc_num_sms() ? g : 3L * sc_num_sms());      // three resident CTAs per SM
  const size_t smem = d.dx_colsum ? (size_t)8 * 3 * d.D * sizeof(float) : (d.dgamma ? (size_t)8 * 2 * d.D * sizeof(float) : 0);
#define SC_LN_CASE(NV_)                                                                                      \
  {                                                                                                          \
    static sc_device_once once;                                                                              \
    if (once.first()) { cudaFuncSetAttribute(ln_bwd_kernel<TDY, TX, TDX, NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304); once.done(); } \
    ln_bwd_kernel<TDY, TX, TDX, NV_><<<grid, 256, smem, st>>>(d);                                           \
  }
  switch (nv) {
    case 1: SC_LN_CASE(1); break;
    case 2: SC_LN_CASE(2); break;
    case 3: SC_LN_CASE(3); break;
    case 4: SC_LN_CASE(4); break;
    case 5: case 6: SC_LN_CASE(6); break;
    case 7: case 8: SC_LN_CASE(8); break;
    default: sc_set_error("sc_layernorm_bwd: D=%d > 1024 unsupported", d.D); return SC_ERR_UNSUPPORTED;
  }
#undef SC_LN_CASE
  SC_LAUNCH_CHECK();
  return SC_OK;
}

}  // namespace

extern void sc_count_launch(int n);

extern "C" int sc_layernorm_fwd(const sc_ln_desc* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(d && d->x && d->y && d->gamma && d->beta, "sc_layernorm_fwd: null pointer");
  SC_CHECK_ARG(d->D % 4 == 0 && d->D > 0, "sc_layernorm_fwd: D=%d must be a multiple of 4", d->D);
  if (d->rows <= 0) return SC_OK;
  sc_count_launch(1);
  if (d->x_dtype == SC_F32 && d->y_dtype == SC_F32) return launch_fwd<float, float>(*d, st);
  if (d->x_dtype == SC_F32 && d->y_dtype == SC_BF16) return launch_fwd<float, bf16>(*d, st);
  if (d->x_dtype == SC_BF16 && d->y_dtype == SC_BF16) return launch_fwd<bf16, bf16>(*d, st);
  if (d->x_dtype == SC_BF16 && d->y_dtype == SC_F32) return launch_fwd<bf16, float>(*d, st);
  sc_set_error("sc_layernorm_fwd: bad dtypes");
  return SC_ERR_INVALID;
}

extern "C" int sc_layernorm_bwd(const sc_ln_bwd_desc* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(d && d->x && d->dy && d->gamma && d->mean && d->rstd, "sc_layernorm_bwd: null pointer");
  SC_CHECK_ARG(d->D % 4 == 0 && d->D > 0, "sc_layernorm_bwd: D=%d must be a multiple of 4", d->D);
  SC_CHECK_ARG((d->dgamma == nullptr) == (d->dbeta == nullptr), "sc_layernorm_bwd: dgamma/dbeta must come together");
  SC_CHECK_ARG(!d->dx_colsum || d->dx, "sc_layernorm_bwd: dx_colsum needs dx");
  if (d->rows <= 0) return SC_OK;
  sc_count_launch(1);
  const int key = d->dy_dtype * 4 + d->x_dtype * 2 + d->dx_dtype;
  static const bool old_kernel = getenv("SC_LN_BWD_SMEM") != nullptr;      // A/B switch: shared-memory accumulator version
  static const bool reg_kernel = getenv("SC_LN_BWD_REGS") != nullptr;      // A/B switch: whole row in registers, no prefetch ring
  const bool vec8 = d->D % 8 == 0 && d->D <= 1024 && ((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->dy & 15) == 0 &&
                    ((uintptr_t)d->dx & 15) == 0 && ((uintptr_t)d->gamma & 15) == 0;
  // the prefetch ring pays off on long rows and many of them (measured: D = 768 x 50176 rows 5.0 vs 4.0 TB/s; D = 512 x 19712
  // rows 3.7 vs 4.6 TB/s: too few rows per warp to amortise the prologue / final reduction)
  static const int tma_min_d = [] { const char* e = getenv("SC_LN_TMA_MIN_D"); return e ? atoi(e) : 640; }();
  if (!old_kernel && !reg_kernel && vec8 && d->D >= tma_min_d && d->rows >= 16384) {
    int rc = SC_ERR_UNSUPPORTED;
    switch (key) {
      case 0: rc = launch_bwd_tma<float, float, float>(*d, st); break;
      case 1: rc = launch_bwd_tma<float, float, bf16>(*d, st); break;
      case 2: rc = launch_bwd_tma<float, bf16, float>(*d, st); break;
      case 3: rc = launch_bwd_tma<float, bf16, bf16>(*d, st); break;
      case 4: rc = launch_bwd_tma<bf16, float, float>(*d, st); break;
      case 5: rc = launch_bwd_tma<bf16, float, bf16>(*d, st); break;
      case 6: rc = launch_bwd_tma<bf16, bf16, float>(*d, st); break;
      case 7: rc = launch_bwd_tma<bf16, bf16, bf16>(*d, st); break;
    }
    if (rc != SC_ERR_UNSUPPORTED) return rc;
  }
  if (!old_kernel && vec8) {
    switch (key) {
      case 0: return launch_bwd8<float, float, float>(*d, st);
      case 1: return launch_bwd8<float, float, bf16>(*d, st);
      case 2: return launch_bwd8<float, bf16, float>(*d, st);
      case 3: return launch_bwd8<float, bf16, bf16>(*d, st);
      case 4: return launch_bwd8<bf16, float, float>(*d, st);
      case 5: return launch_bwd8<bf16, float, bf16>(*d, st);
      case 6: return launch_bwd8<bf16, bf16, float>(*d, st);
      case 7: return launch_bwd8<bf16, bf16, bf16>(*d, st);
    }
  }
  switch (key) {
    case 0: return launch_bwd<float, float, float>(*d, st);
    case 1: return launch_bwd<float, float, bf16>(*d, st);
    case 2: return launch_bwd<float, bf16, float>(*d, st);
    case 3: return launch_bwd<float, bf16, bf16>(*d, st);
    case 4: return launch_bwd<bf16, float, float>(*d, st);
    case 5: return launch_bwd<bf16, float, bf16>(*d, st);
    case 6: return launch_bwd<bf16, bf16, float>(*d, st);
    case 7: return launch_bwd<bf16, bf16, bf16>(*d, st);
  }
  sc_set_error("sc_layernorm_bwd: bad dtypes");
  return SC_ERR_INVALID;
}
