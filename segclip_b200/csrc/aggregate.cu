// Learnable-centre patch aggregation (SemanticLearnerModule, modules/module_seg_vit.py:304-312) and
// ReconstructLayer (:333-345): assignment logits, Gumbel-softmax over the centres, hard (arg-max)
// assignment with straight-through gradient, per-centre weighted mean, and their backward passes.
// All fp32 (the discrete assignment must not see bf16 noise beyond the k_conv GEMM).
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

int sc_assign_bwd(const sc_assign_bwd_desc* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SC_CHECK_ARG(a && a->d_agg && a->agg && a->v && a->idx && a->count && a->y_soft && a->d_logits && a->d_v && a->qf &&
                   a->k && a->d_k && a->d_qf, "sc_assign_bwd: null pointer");
  SC_CHECK_ARG(a->G == G, "sc_assign_bwd: G=%d (only 8 centres supported)", a->G);
  const int B = a->B, L = a->L, D = a->D;
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
