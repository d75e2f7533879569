// Learnable-centre patch aggregation (SemanticLearnerModule, modules/module_seg_vit.py:304-312) and
// ReconstructLayer (:333-345): assignment logits, Gumbel-softmax over the centres, hard (arg-max)
// assignment with straight-through gradient, per-centre weighted mean, and their backward passes.
// All fp32 (the discrete assignment must not see bf16 noise beyond the k_conv GEMM).
#include <stdlib.h>

#include "common.cuh"

extern void sc_count_launch(int n);

namespace {

constexpr int G = 8;  // number of centres (group_num, module_seg_vit.py:349)

// Gumbel(0,1) noise as torch.distributions.Gumbel draws it from u ~ torch.rand:
// Uniform(tiny, 1-eps).sample -> -log(-log(.))  (module_seg_vit.py:223-226)
SC_DEVINL float gumbel_from_uniform(float u) {
  const float tiny = 1.17549435e-38f, eps = 1.1920929e-07f;
  const float base = tiny + u * ((1.0f - eps) - tiny);
  return -logf(-logf(base));
}

// one warp per patch: 8 dot products of length D against the sample's centre queries (smem)
template <typename TK>
__global__ void __launch_bounds__(256) assign_fwd_kernel(sc_assign_desc a) {
  extern __shared__ float sq[];  // [G][D]
  const int b = blockIdx.y, D = a.D, L = a.L;
  for (int i = threadIdx.x; i < G * D; i += 256) sq[i] = a.qf[(long)b * G * D + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = blockIdx.x * 8 + warp;
  if (l >= L) return;
  const TK* k = (const TK*)a.k + ((long)b * L + l) * D;
  float acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float kv = to_f32(k[d]);
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = fmaf(sq[g * D + d], kv, acc[g]);
  }
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = warp_sum(acc[g]);
  if (lane != 0) return;
  float z[G], mz = -INFINITY, ml = -INFINITY;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    // training: (logits + Gumbel) / tau; inference (u == NULL): plain logits (module_seg_vit.py:228-231)
    z[g] = a.u ? (acc[g] + gumbel_from_uniform(a.u[((long)b * G + g) * L + l])) / a.tau : acc[g];
    mz = fmaxf(mz, z[g]);
    ml = fmaxf(ml, acc[g]);
  }
  float sz = 0.f, sl = 0.f;
  float ez[G], el[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ez[g] = expf(z[g] - mz);
    el[g] = expf(acc[g] - ml);
    sz += ez[g];
    sl += el[g];
  }
  int arg = 0;
  float best = -1.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float y = ez[g] / sz;
    a.y_soft[((long)b * G + g) * L + l] = y;
    if (a.soft) a.soft[((long)b * G + g) * L + l] = el[g] / sl;
    if (a.logits) a.logits[((long)b * G + g) * L + l] = acc[g];
    if (y > best) { best = y; arg = g; }   // first maximum, like Tensor.max(dim)
  }
  if (a.forced_idx) arg = a.forced_idx[(long)b * L + l];
  a.idx[(long)b * L + l] = arg;
  atomicAdd(a.count + b * G + arg, 1.0f);
}

// agg[b,g,:] = sum_{l: idx=g} v[b,l,:] / max(count,1);  sum_out = qf + agg
template <typename TV>
__global__ void __launch_bounds__(128) aggregate_fwd_kernel(const TV* __restrict__ v, const int* __restrict__ idx,
                                                             const float* __restrict__ count, const float* __restrict__ qf,
                                                             float* __restrict__ agg, float* __restrict__ sum_out, int L, int D) {
  extern __shared__ int sidx[];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < L; i += 128) sidx[i] = idx[(long)b * L + i];
  __syncthreads();
  const int d = blockIdx.x * 128 + threadIdx.x;
  if (d >= D) return;
  float acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = 0.f;
  for (int l = 0; l < L; ++l) {
    const float x = to_f32(v[((long)b * L + l) * D + d]);
    const int c = sidx[l];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] += (c == g) ? x : 0.f;
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float m = acc[g] / fmaxf(count[b * G + g], 1.0f);
    const long o = ((long)b * G + g) * D + d;
    agg[o] = m;
    sum_out[o] = qf[o] + m;
  }
}

// Backward of the hard assignment: per patch, d hard -> d logits (through y_soft only), and d v.
template <typename TV>
__global__ void __launch_bounds__(256) assign_bwd_kernel(sc_assign_bwd_desc a) {
  extern __shared__ float sm[];
  const int b = blockIdx.y, D = a.D, L = a.L;
  float* sdo = sm;           // [G][D] d agg
  float* st = sm + G * D;    // [G] dAgg_g . agg_g
  for (int i = threadIdx.x; i < G * D; i += 256) sdo[i] = a.d_agg[(long)b * G * D + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // warp g computes t_g
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(sdo[warp * D + d], a.agg[((long)b * G + warp) * D + d], s);
    s = warp_sum(s);
    if (lane == 0) st[warp] = s;
  }
  __syncthreads();
  const int l = blockIdx.x * 8 + warp;
  if (l >= L) return;
  const TV* v = (const TV*)a.v + ((long)b * L + l) * D;
  TV* dv = (TV*)a.d_v + ((long)b * L + l) * D;
  const int c = a.idx[(long)b * L + l];
  const float inv_c = 1.0f / fmaxf(a.count[b * G + c], 1.0f);
  float acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float x = to_f32(v[d]);
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = fmaf(sdo[g * D + d], x, acc[g]);
    dv[d] = from_f32<TV>(sdo[c * D + d] * inv_c);
  }
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = warp_sum(acc[g]);
  if (lane != 0) return;
  float dy[G], y[G], dot = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float cnt = a.count[b * G + g];
    // out_g = S_g / max(cnt,1); clamp_min passes gradient for cnt >= 1 (module_seg_vit.py:310)
    float dh = (acc[g] - (cnt >= 1.0f ? st[g] : 0.f)) / fmaxf(cnt, 1.0f);
    if (a.d_hard_extra) dh += a.d_hard_extra[((long)b * G + g) * L + l];
    y[g] = a.y_soft[((long)b * G + g) * L + l];
    dy[g] = dh;
    dot += y[g] * dh;
  }
#pragma unroll
  for (int g = 0; g < G; ++g) a.d_logits[((long)b * G + g) * L + l] = y[g] * (dy[g] - dot) / a.tau;
}

// dk[b,l,:] = sum_g dlogit[b,g,l] qf[b,g,:]
template <typename TK>
__global__ void __launch_bounds__(128) assign_bwd_dk_kernel(const float* __restrict__ dlog, const float* __restrict__ qf,
                                                             TK* __restrict__ dk, int L, int D) {
  const int l = blockIdx.x, b = blockIdx.y;
  float w[G];
#pragma unroll
  for (int g = 0; g < G; ++g) w[g] = dlog[((long)b * G + g) * L + l];
  for (int d = threadIdx.x; d < D; d += 128) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) s = fmaf(w[g], qf[((long)b * G + g) * D + d], s);
    dk[((long)b * L + l) * D + d] = from_f32<TK>(s);
  }
}
// dqf[b,g,:] = base[b,g,:] + sum_l dlogit[b,g,l] k[b,l,:]
template <typename TK>
__global__ void __launch_bounds__(128) assign_bwd_dq_kernel(const float* __restrict__ dlog, const TK* __restrict__ k,
                                                             const float* __restrict__ base, float* __restrict__ dqf, int L, int D) {
  extern __shared__ float sw[];  // [G][L]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < G * L; i += 128) sw[i] = dlog[(long)b * G * L + i];
  __syncthreads();
  const int d = blockIdx.x * 128 + threadIdx.x;
  if (d >= D) return;
  float acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = base ? base[((long)b * G + g) * D + d] : 0.f;
  for (int l = 0; l < L; ++l) {
    const float x = to_f32(k[((long)b * L + l) * D + d]);
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = fmaf(sw[g * L + l], x, acc[g]);
  }
#pragma unroll
  for (int g = 0; g < G; ++g) dqf[((long)b * G + g) * D + d] = acc[g];
}

// ---------------------------------------------------------------- fused aggregation (one CTA per sample)
// The patch -> centre assignment softmax, the hard arg-max and the per-centre weighted sum as ONE coalesced kernel
// (module_seg_vit.py:304-312): a sample's [L, D] key / value tiles are each read exactly once with 16-byte row accesses.
//   phase 1 (warp per patch)      : 8 dots <qf_g, k_l> against the centre queries in shared memory, Gumbel softmax, arg-max;
//                                   the patch's centre and the per-centre counts stay in shared memory
//   phase 2 (thread per 4 columns): agg[g, cols] = sum_{l: idx_l = g} v[l, cols] / max(count_g, 1); sum_out = qf + agg
// All 2 x B CTAs of the benchmark batch are resident at once (<= 64 registers, ~32 KB of shared memory).
constexpr int FA_THREADS = 512;
constexpr int FA_ROWS = 4;        // patches per warp and trip in the warp-per-patch phases

template <typename T> SC_DEVINL float4 ld4(const T* p);
template <> SC_DEVINL float4 ld4<float>(const float* p) { return *(const float4*)p; }
template <> SC_DEVINL float4 ld4<bf16>(const bf16* p) {
  const uint2 u = *(const uint2*)p;
  const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&u.x), b = __bfloat1622float2(*(const __nv_bfloat162*)&u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> SC_DEVINL void st4(T* p, const float4& v);
template <> SC_DEVINL void st4<float>(float* p, const float4& v) { *(float4*)p = v; }
template <> SC_DEVINL void st4<bf16>(bf16* p, const float4& v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *(uint32_t*)&a;
  u.y = *(uint32_t*)&b;
  *(uint2*)p = u;
}

// the Gumbel-softmax / arg-max of one patch from its 8 logits (lane 0 of the patch's warp); same arithmetic as assign_fwd_kernel
SC_DEVINL int assign_one(const sc_assign_desc& a, int b, int l, const float (&acc)[G]) {
  const int L = a.L;
  float z[G], mz = -INFINITY, ml = -INFINITY;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    z[g] = a.u ? (acc[g] + gumbel_from_uniform(a.u[((long)b * G + g) * L + l])) / a.tau : acc[g];
    mz = fmaxf(mz, z[g]);
    ml = fmaxf(ml, acc[g]);
  }
  float sz = 0.f, sl = 0.f, ez[G], el[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ez[g] = expf(z[g] - mz);
    el[g] = expf(acc[g] - ml);
    sz += ez[g];
    sl += el[g];
  }
  int arg = 0;
  float best = -1.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float y = ez[g] / sz;
    a.y_soft[((long)b * G + g) * L + l] = y;
    if (a.soft) a.soft[((long)b * G + g) * L + l] = el[g] / sl;
    if (a.logits) a.logits[((long)b * G + g) * L + l] = acc[g];
    if (y > best) { best = y; arg = g; }   // first maximum, like Tensor.max(dim)
  }
  if (a.forced_idx) arg = a.forced_idx[(long)b * L + l];
  a.idx[(long)b * L + l] = arg;
  return arg;
}

template <typename TK, typename TV>
__global__ void __launch_bounds__(FA_THREADS, 2)
assign_aggregate_fwd_kernel(sc_assign_desc a, const TV* __restrict__ v, float* __restrict__ agg, float* __restrict__ sum_out) {
  extern __shared__ __align__(16) float fsm[];
  const int b = blockIdx.x, D = a.D, L = a.L, D4 = D >> 2;
  float* sq = fsm;                       // [G][D] centre queries
  float* ssum = sq + G * D;              // [G][D] per-centre sums (phase 2)
  float* scount = ssum + G * D;          // [G]
  int* sidx = (int*)(scount + G);        // [L]
  for (int i = threadIdx.x; i < G * D4; i += FA_THREADS) {
    ((float4*)sq)[i] = ((const float4*)(a.qf + (long)b * G * D))[i];
    ((float4*)ssum)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x < G) scount[threadIdx.x] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- phase 1: FA_ROWS patches per warp and trip, so that every 16-byte read of a centre query from shared memory feeds
  // FA_ROWS dot products (one patch per trip made the kernel shared-memory-bandwidth bound: 24 KB of queries per patch)
  for (int l0 = warp; l0 < L; l0 += (FA_THREADS / 32) * FA_ROWS) {
    float acc[FA_ROWS][G];
#pragma unroll
    for (int r = 0; r < FA_ROWS; ++r)
#pragma unroll
      for (int g = 0; g < G; ++g) acc[r][g] = 0.f;
    const TK* k0 = (const TK*)a.k + ((long)b * L + l0) * D;
    for (int i4 = lane; i4 < D4; i4 += 32) {
      float4 kv[FA_ROWS];
#pragma unroll
      for (int r = 0; r < FA_ROWS; ++r)
        kv[r] = (l0 + r * (FA_THREADS / 32) < L) ? ld4<TK>(k0 + (long)r * (FA_THREADS / 32) * D + 4 * i4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 q = ((const float4*)(sq + g * D))[i4];
#pragma unroll
        for (int r = 0; r < FA_ROWS; ++r)
          acc[r][g] = fmaf(q.x, kv[r].x, fmaf(q.y, kv[r].y, fmaf(q.z, kv[r].z, fmaf(q.w, kv[r].w, acc[r][g]))));
      }
    }
    float mine[G];                               // lane r finishes patch r of the trip
#pragma unroll
    for (int r = 0; r < FA_ROWS; ++r)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float t = warp_sum(acc[r][g]);
        if (lane == r) mine[g] = t;
      }
    const int l = l0 + lane * (FA_THREADS / 32);
    if (lane < FA_ROWS && l < L) {
      const int arg = assign_one(a, b, l, mine);
      sidx[l] = arg;
      atomicAdd(scount + arg, 1.0f);
    }
  }
  __syncthreads();
  // ---- phase 2
  const int rgroups = FA_THREADS / D4;           // row groups working side by side (host guarantees >= 1)
  const int rg = threadIdx.x / D4, c4 = threadIdx.x - rg * D4;
  if (rg < rgroups) {
    float4 s[G];
#pragma unroll
    for (int g = 0; g < G; ++g) s[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    const TV* vp = v + (long)b * L * D + 4 * c4;
#pragma unroll 4
    for (int l = rg; l < L; l += rgroups) {
      const float4 x = ld4<TV>(vp + (long)l * D);
      const int c = sidx[l];
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (c == g) { s[g].x += x.x; s[g].y += x.y; s[g].z += x.z; s[g].w += x.w; }
    }
    if (rgroups == 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) ((float4*)(ssum + g * D))[c4] = s[g];
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float* o = ssum + g * D + 4 * c4;
        atomicAdd(o, s[g].x); atomicAdd(o + 1, s[g].y); atomicAdd(o + 2, s[g].z); atomicAdd(o + 3, s[g].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * D4; i += FA_THREADS) {
    const int g = i / D4;
    const float inv = 1.0f / fmaxf(scount[g], 1.0f);
    float4 m = ((const float4*)ssum)[i];
    const float4 q = ((const float4*)sq)[i];
    m.x *= inv; m.y *= inv; m.z *= inv; m.w *= inv;
    ((float4*)(agg + (long)b * G * D))[i] = m;
    ((float4*)(sum_out + (long)b * G * D))[i] = make_float4(q.x + m.x, q.y + m.y, q.z + m.z, q.w + m.w);
  }
  if (threadIdx.x < G) a.count[b * G + threadIdx.x] = scount[threadIdx.x];
}

// Backward of the fused kernel (sc_assign_bwd contract), one CTA per sample, every [L, D] tile touched once:
//   phase B (warp per patch)      : 8 dots <d_agg_g, v_l> -> d hard -> d logits (kept in shared memory); d v_l = d_agg[idx_l] / count
//   phase C (thread per 4 columns): d k_l = sum_g d logits[g,l] qf_g (written) and d qf_g += sum_l d logits[g,l] k_l (k read) in one sweep
template <typename TK, typename TV>
__global__ void __launch_bounds__(FA_THREADS, 2) assign_bwd_fused_kernel(sc_assign_bwd_desc a) {
  extern __shared__ __align__(16) float fsm[];
  const int b = blockIdx.x, D = a.D, L = a.L, D4 = D >> 2;
  float* sa = fsm;                       // [G][D]: d agg (phase B), then qf (phase C)
  float* sacc = sa + G * D;              // [G][D]: d qf accumulators (phase C)
  float* st = sacc + G * D;              // [G]  <d agg_g, agg_g>
  float* sinv = st + G;                  // [G]  1 / max(count, 1)
  float* sdl = sinv + G;                 // [G][L] d logits
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < G * D4; i += FA_THREADS) ((float4*)sa)[i] = ((const float4*)(a.d_agg + (long)b * G * D))[i];
  if (threadIdx.x < G) sinv[threadIdx.x] = 1.0f / fmaxf(a.count[b * G + threadIdx.x], 1.0f);
  __syncthreads();
  if (warp < G) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(sa[warp * D + d], a.agg[((long)b * G + warp) * D + d], s);
    s = warp_sum(s);
    if (lane == 0) st[warp] = s;
  }
  __syncthreads();
  // ---- phase B (FA_ROWS patches per warp and trip, see the forward kernel)
  for (int l0 = warp; l0 < L; l0 += (FA_THREADS / 32) * FA_ROWS) {
    float acc[FA_ROWS][G];
    int cr[FA_ROWS];
#pragma unroll
    for (int r = 0; r < FA_ROWS; ++r) {
      const int l = l0 + r * (FA_THREADS / 32);
      cr[r] = l < L ? a.idx[(long)b * L + l] : -1;
#pragma unroll
      for (int g = 0; g < G; ++g) acc[r][g] = 0.f;
    }
    const TV* v0 = (const TV*)a.v + ((long)b * L + l0) * D;
    TV* dv0 = (TV*)a.d_v + ((long)b * L + l0) * D;
    for (int i4 = lane; i4 < D4; i4 += 32) {
      float4 x[FA_ROWS];
#pragma unroll
      for (int r = 0; r < FA_ROWS; ++r)
        x[r] = cr[r] >= 0 ? ld4<TV>(v0 + (long)r * (FA_THREADS / 32) * D + 4 * i4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 q = ((const float4*)(sa + g * D))[i4];
#pragma unroll
        for (int r = 0; r < FA_ROWS; ++r)
          acc[r][g] = fmaf(q.x, x[r].x, fmaf(q.y, x[r].y, fmaf(q.z, x[r].z, fmaf(q.w, x[r].w, acc[r][g]))));
      }
#pragma unroll
      for (int r = 0; r < FA_ROWS; ++r)
        if (cr[r] >= 0) {
          const float inv_c = sinv[cr[r]];
          const float4 dc = ((const float4*)(sa + cr[r] * D))[i4];
          st4<TV>(dv0 + (long)r * (FA_THREADS / 32) * D + 4 * i4, make_float4(dc.x * inv_c, dc.y * inv_c, dc.z * inv_c, dc.w * inv_c));
        }
    }
    float mine[G];
#pragma unroll
    for (int r = 0; r < FA_ROWS; ++r)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float t = warp_sum(acc[r][g]);
        if (lane == r) mine[g] = t;
      }
    const int l = l0 + lane * (FA_THREADS / 32);
    if (lane < FA_ROWS && l < L) {
      float dy[G], y[G], dot = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float cnt = a.count[b * G + g];
        // out_g = S_g / max(cnt,1); clamp_min passes gradient for cnt >= 1 (module_seg_vit.py:310)
        float dh = (mine[g] - (cnt >= 1.0f ? st[g] : 0.f)) / fmaxf(cnt, 1.0f);
        if (a.d_hard_extra) dh += a.d_hard_extra[((long)b * G + g) * L + l];
        y[g] = a.y_soft[((long)b * G + g) * L + l];
        dy[g] = dh;
        dot += y[g] * dh;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float dl = y[g] * (dy[g] - dot) / a.tau;
        sdl[g * L + l] = dl;
        a.d_logits[((long)b * G + g) * L + l] = dl;
      }
    }
  }
  __syncthreads();
  // ---- phase C
  for (int i = threadIdx.x; i < G * D4; i += FA_THREADS) {
    ((float4*)sa)[i] = ((const float4*)(a.qf + (long)b * G * D))[i];
    ((float4*)sacc)[i] = a.d_qf_base ? ((const float4*)(a.d_qf_base + (long)b * G * D))[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int rgroups = FA_THREADS / D4;
  const int rg = threadIdx.x / D4, c4 = threadIdx.x - rg * D4;
  if (rg < rgroups) {
    {   // sweep 1 (stores only): d k_l = sum_g d logits[g,l] qf_g, the thread's slice of the eight queries in registers
      float4 q[G];
#pragma unroll
      for (int g = 0; g < G; ++g) q[g] = ((const float4*)(sa + g * D))[c4];
      TK* dkp = (TK*)a.d_k + (long)b * L * D + 4 * c4;
#pragma unroll 2
      for (int l = rg; l < L; l += rgroups) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float w = sdl[g * L + l];
          o.x = fmaf(w, q[g].x, o.x); o.y = fmaf(w, q[g].y, o.y); o.z = fmaf(w, q[g].z, o.z); o.w = fmaf(w, q[g].w, o.w);
        }
        st4<TK>(dkp + (long)l * D, o);
      }
    }
    // sweep 2 (loads only): d qf_g += sum_l d logits[g,l] k_l, four key rows requested per trip
    float4 s[G];
#pragma unroll
    for (int g = 0; g < G; ++g) s[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    const TK* kp = (const TK*)a.k + (long)b * L * D + 4 * c4;
#pragma unroll 1
    for (int l = rg; l < L; l += 4 * rgroups) {
      float4 x[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) x[r] = (l + r * rgroups < L) ? ld4<TK>(kp + (long)(l + r * rgroups) * D) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int lr = min(l + r * rgroups, L - 1);            // (rows past L carry x = 0)
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float w = sdl[g * L + lr];
          s[g].x = fmaf(w, x[r].x, s[g].x); s[g].y = fmaf(w, x[r].y, s[g].y); s[g].z = fmaf(w, x[r].z, s[g].z); s[g].w = fmaf(w, x[r].w, s[g].w);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float* o = sacc + g * D + 4 * c4;
      atomicAdd(o, s[g].x); atomicAdd(o + 1, s[g].y); atomicAdd(o + 2, s[g].z); atomicAdd(o + 3, s[g].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * D4; i += FA_THREADS) ((float4*)(a.d_qf + (long)b * G * D))[i] = ((const float4*)sacc)[i];
}

static bool fused_aggregation_ok(int D, int L, const void* p0, const void* p1) {
  static const bool off = getenv("SC_AGG_UNFUSED") != nullptr;      // A/B switch: the separate assign / aggregate kernels
  return !off && D % 4 == 0 && D / 4 <= FA_THREADS && D <= 1536 && L <= 4096 && (((uintptr_t)p0 | (uintptr_t)p1) & 15) == 0;
}

// ---------------------------------------------------------------- ReconstructLayer
// pre[b,m,:] = sum_g' (W[g', idx[b,m]] + bias[g']) * sx[b,g',:] ; out = QuickGELU(pre)
__global__ void __launch_bounds__(128) reconstruct_fwd_kernel(const float* __restrict__ sx, const int* __restrict__ idx,
                                                               const float* __restrict__ W, const float* __restrict__ bias,
                                                               float* __restrict__ pre, float* __restrict__ out, int M, int D) {
  const int m = blockIdx.x, b = blockIdx.y;
  const int c = idx[(long)b * M + m];
  float a[G];
#pragma unroll
  for (int g = 0; g < G; ++g) a[g] = W[g * G + c] + bias[g];
  for (int d = threadIdx.x; d < D; d += 128) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) s = fmaf(a[g], sx[((long)b * G + g) * D + d], s);
    const long o = ((long)b * M + m) * D + d;
    pre[o] = s;
    out[o] = act_fwd(s, SC_ACT_QUICKGELU);
  }
}
// per sample: dA'[m,g'] = sum_h dpre[m,h] sx[g',h]; d hard[g,m] = sum_g' dA'[m,g'] W[g',g];
// dW[g',g] += sum_{m: idx=g} dA'[m,g']; dbias[g'] += sum_m dA'[m,g']
__global__ void __launch_bounds__(256) reconstruct_bwd_a_kernel(const float* __restrict__ d_out, const float* __restrict__ pre,
                                                                 const float* __restrict__ sx, const int* __restrict__ idx,
                                                                 const float* __restrict__ W, float* __restrict__ d_hard,
                                                                 float* __restrict__ dW, float* __restrict__ dbias, int M, int D) {
  extern __shared__ float sm[];
  float* ssx = sm;              // [G][D]
  float* sdW = sm + G * D;      // [G*G]
  float* sdb = sdW + G * G;     // [G]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < G * D; i += 256) ssx[i] = sx[(long)b * G * D + i];
  if (threadIdx.x < G * G + G) sdW[threadIdx.x] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < M; m += 8) {
    const long row = ((long)b * M + m) * D;
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float dp = d_out[row + d] * act_grad(pre[row + d], SC_ACT_QUICKGELU);
#pragma unroll
      for (int g = 0; g < G; ++g) acc[g] = fmaf(dp, ssx[g * D + d], acc[g]);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = warp_sum(acc[g]);
    if (lane == 0) {
      const int c = idx[(long)b * M + m];
#pragma unroll
      for (int gp = 0; gp < G; ++gp) {
        atomicAdd(&sdW[gp * G + c], acc[gp]);
        atomicAdd(&sdb[gp], acc[gp]);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float s = 0.f;
#pragma unroll
        for (int gp = 0; gp < G; ++gp) s = fmaf(acc[gp], W[gp * G + g], s);
        d_hard[((long)b * G + g) * M + m] = s;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < G * G) atomicAdd(dW + threadIdx.x, sdW[threadIdx.x]);
  else if (threadIdx.x < G * G + G) atomicAdd(dbias + threadIdx.x - G * G, sdb[threadIdx.x - G * G]);
}
// dsx[b,g',:] = sum_m (W[g', idx[m]] + bias[g']) dpre[b,m,:]
__global__ void __launch_bounds__(128) reconstruct_bwd_sx_kernel(const float* __restrict__ d_out, const float* __restrict__ pre,
                                                                  const int* __restrict__ idx, const float* __restrict__ W,
                                                                  const float* __restrict__ bias, float* __restrict__ dsx, int M, int D) {
  extern __shared__ float sa[];  // [M][G]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < M * G; i += 128) {
    const int m = i / G, g = i % G;
    sa[i] = W[g * G + idx[(long)b * M + m]] + bias[g];
  }
  __syncthreads();
  const int d = blockIdx.x * 128 + threadIdx.x;
  if (d >= D) return;
  float acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = 0.f;
  for (int m = 0; m < M; ++m) {
    const long o = ((long)b * M + m) * D + d;
    const float dp = d_out[o] * act_grad(pre[o], SC_ACT_QUICKGELU);
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = fmaf(sa[m * G + g], dp, acc[g]);
  }
#pragma unroll
  for (int g = 0; g < G; ++g) dsx[((long)b * G + g) * D + d] = acc[g];
}

}  // namespace

extern "C" {

int sc_assign_fwd(const sc_assign_desc* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(a && a->qf && a->k && a->y_soft && a->idx && a->count, "sc_assign_fwd: null pointer");
  SC_CHECK_ARG(a->G == G, "sc_assign_fwd: G=%d (only 8 centres supported)", a->G);
  const size_t smem = sizeof(float) * G * a->D;
  dim3 grid(ceil_div(a->L, 8), a->B);
  sc_count_launch(1);
  if (a->k_dtype == SC_F32) assign_fwd_kernel<float><<<grid, 256, smem, st>>>(*a);
  else assign_fwd_kernel<bf16><<<grid, 256, smem, st>>>(*a);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_aggregate_fwd(const void* v, int v_dtype, const int32_t* idx, const float* count, const float* qf, float* agg,
                     float* sum_out, int B, int L, int D, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(v && idx && count && qf && agg && sum_out, "sc_aggregate_fwd: null pointer");
  dim3 grid(ceil_div(D, 128), B);
  sc_count_launch(1);
  if (v_dtype == SC_F32)
    aggregate_fwd_kernel<float><<<grid, 128, L * sizeof(int), st>>>((const float*)v, idx, count, qf, agg, sum_out, L, D);
  else
    aggregate_fwd_kernel<bf16><<<grid, 128, L * sizeof(int), st>>>((const bf16*)v, idx, count, qf, agg, sum_out, L, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_assign_aggregate_fwd(const sc_assign_desc* a, const void* v, int v_dtype, float* agg, float* sum_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(a && a->qf && a->k && a->y_soft && a->idx && a->count && v && agg && sum_out, "sc_assign_aggregate_fwd: null pointer");
  SC_CHECK_ARG(a->G == G, "sc_assign_aggregate_fwd: G=%d (only 8 centres supported)", a->G);
  if (!fused_aggregation_ok(a->D, a->L, a->k, v) || (((uintptr_t)a->qf | (uintptr_t)agg | (uintptr_t)sum_out) & 15) != 0) {
    // shapes the one-CTA-per-sample kernel does not take: the two-kernel path (count must be zero on entry)
    SC_CUDA(cudaMemsetAsync(a->count, 0, sizeof(float) * a->B * G, st));
    int rc = sc_assign_fwd(a, stream);
    if (rc) return rc;
    return sc_aggregate_fwd(v, v_dtype, a->idx, a->count, a->qf, agg, sum_out, a->B, a->L, a->D, stream);
  }
  const size_t smem = sizeof(float) * (2 * G * a->D + G) + sizeof(int) * a->L;
  sc_count_launch(1);
#define SC_FA_CASE(TK_, TV_)                                                                                              \
  {                                                                                                                       \
    static sc_device_once once;                                                                                           \
    if (once.first()) { SC_CUDA(cudaFuncSetAttribute(assign_aggregate_fwd_kernel<TK_, TV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024)); once.done(); } \
    assign_aggregate_fwd_kernel<TK_, TV_><<<a->B, FA_THREADS, smem, st>>>(*a, (const TV_*)v, agg, sum_out);               \
  }
  if (a->k_dtype == SC_F32 && v_dtype == SC_F32) SC_FA_CASE(float, float)
  else if (a->k_dtype == SC_F32) SC_FA_CASE(float, bf16)
  else if (v_dtype == SC_F32) SC_FA_CASE(bf16, float)
  else SC_FA_CASE(bf16, bf16)
#undef SC_FA_CASE
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_assign_bwd(const sc_assign_bwd_desc* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(a && a->d_agg && a->agg && a->v && a->idx && a->count && a->y_soft && a->d_logits && a->d_v && a->qf &&
                   a->k && a->d_k && a->d_qf, "sc_assign_bwd: null pointer");
  SC_CHECK_ARG(a->G == G, "sc_assign_bwd: G=%d (only 8 centres supported)", a->G);
  const int B = a->B, L = a->L, D = a->D;
  if (fused_aggregation_ok(D, L, a->k, a->v) &&
      (((uintptr_t)a->d_k | (uintptr_t)a->d_v | (uintptr_t)a->d_agg | (uintptr_t)a->qf | (uintptr_t)a->d_qf | (uintptr_t)a->d_qf_base) & 15) == 0) {
    const size_t smem = sizeof(float) * (2 * G * D + 2 * G + G * L);
    sc_count_launch(1);
#define SC_FB_CASE(TK_, TV_)                                                                                              \
  {                                                                                                                       \
    static sc_device_once once;                                                                                           \
    if (once.first()) { SC_CUDA(cudaFuncSetAttribute(assign_bwd_fused_kernel<TK_, TV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); once.done(); } \
    assign_bwd_fused_kernel<TK_, TV_><<<B, FA_THREADS, smem, st>>>(*a);                                                   \
  }
    if (a->k_dtype == SC_F32 && a->v_dtype == SC_F32) SC_FB_CASE(float, float)
    else if (a->k_dtype == SC_F32) SC_FB_CASE(float, bf16)
    else if (a->v_dtype == SC_F32) SC_FB_CASE(bf16, float)
    else SC_FB_CASE(bf16, bf16)
#undef SC_FB_CASE
    SC_LAUNCH_CHECK();
    return SC_OK;
  }
  sc_count_launch(3);
  {
    dim3 grid(ceil_div(L, 8), B);
    const size_t smem = sizeof(float) * (G * D + G);
    if (a->v_dtype == SC_F32) assign_bwd_kernel<float><<<grid, 256, smem, st>>>(*a);
    else assign_bwd_kernel<bf16><<<grid, 256, smem, st>>>(*a);
  }
  if (a->k_dtype == SC_F32) {
    assign_bwd_dk_kernel<float><<<dim3(L, B), 128, 0, st>>>(a->d_logits, a->qf, (float*)a->d_k, L, D);
    assign_bwd_dq_kernel<float><<<dim3(ceil_div(D, 128), B), 128, sizeof(float) * G * L, st>>>(
        a->d_logits, (const float*)a->k, a->d_qf_base, a->d_qf, L, D);
  } else {
    assign_bwd_dk_kernel<bf16><<<dim3(L, B), 128, 0, st>>>(a->d_logits, a->qf, (bf16*)a->d_k, L, D);
    assign_bwd_dq_kernel<bf16><<<dim3(ceil_div(D, 128), B), 128, sizeof(float) * G * L, st>>>(
        a->d_logits, (const bf16*)a->k, a->d_qf_base, a->d_qf, L, D);
  }
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_reconstruct_fwd(const float* sx, const int32_t* idx, const float* W, const float* bias, float* pre, float* out,
                       int B, int M, int D, void* stream) {
  SC_CHECK_ARG(sx && idx && W && bias && pre && out, "sc_reconstruct_fwd: null pointer");
  sc_count_launch(1);
  reconstruct_fwd_kernel<<<dim3(M, B), 128, 0, (cudaStream_t)stream>>>(sx, idx, W, bias, pre, out, M, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_reconstruct_bwd(const float* d_out, const float* pre, const float* sx, const int32_t* idx, const float* W,
                       const float* bias, float* d_sx, float* d_hard, float* dW, float* dbias, int B, int M, int D,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(d_out && pre && sx && idx && W && bias && d_sx && d_hard && dW && dbias, "sc_reconstruct_bwd: null pointer");
  sc_count_launch(2);
  reconstruct_bwd_a_kernel<<<B, 256, sizeof(float) * (G * D + G * G + G), st>>>(d_out, pre, sx, idx, W, d_hard, dW, dbias, M, D);
  reconstruct_bwd_sx_kernel<<<dim3(ceil_div(D, 128), B), 128, sizeof(float) * M * G, st>>>(d_out, pre, idx, W, bias, d_sx, M, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

}  // extern "C"
