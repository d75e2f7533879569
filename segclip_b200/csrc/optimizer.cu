// Fused multi-tensor optimizer step for the "next" row of SURVEY 8(f): clip_grad_norm_ + AdaptAdamW.step + logit_scale clamp
// (reference main_task_align.py:326-347, modules/optimization_adamw.py:112-174).  Two launches for the whole model instead
// of ~10 tiny kernels per parameter tensor from a Python loop: (1) sum of squared gradients, (2) the update.
// Purely HBM-bound: reads p, g, m, v and writes p, m, v once (28 B per parameter).
#include "common.cuh"

extern void sc_count_launch(int n);

namespace {

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const sc_opt_item* __restrict__ items, int n_items, float* __restrict__ out) {
  __shared__ float red[8];
  int lo = 0, hi = n_items - 1;
  const long blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const sc_opt_item it = items[lo];
  const float* g = (const float*)it.grad;
  const long base = (blk - it.first_block) * 1024 + threadIdx.x * 4;
  float s = 0.f;
  if (g != nullptr) {
    if (base + 3 < it.n && ((uintptr_t)(g + base) & 15) == 0) {
      const float4 v = *(const float4*)(g + base);
      s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long i = base; i < base + 4 && i < it.n; ++i) s += g[i] * g[i];
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(const sc_opt_item* __restrict__ items, int n_items, const float* __restrict__ sqnorm,
                                                     float max_norm, float beta1, float beta2, float eps) {
  int lo = 0, hi = n_items - 1;
  const long blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const sc_opt_item it = items[lo];
  if (it.grad == nullptr) return;
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
  float clip = 1.0f;
  if (max_norm > 0.f) clip = fminf(1.0f, max_norm / (sqrtf(*sqnorm) + 1e-6f));
  float* p = (float*)it.param;
  const float* g = (const float*)it.grad;
  float* m = (float*)it.exp_avg;
  float* v = (float*)it.exp_avg_sq;
  const long base = (blk - it.first_block) * 1024 + threadIdx.x * 4;
  if (base + 3 < it.n && (((uintptr_t)(p + base) | (uintptr_t)(g + base) | (uintptr_t)(m + base) | (uintptr_t)(v + base)) & 15) == 0) {
    float4 p4 = *(float4*)(p + base), m4 = *(float4*)(m + base), v4 = *(float4*)(v + base);
    const float4 g4 = *(const float4*)(g + base);
    float* pp = &p4.x; float* mm = &m4.x; float* vv = &v4.x; const float* gg = &g4.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gi = gg[k] * clip;
      mm[k] = beta1 * mm[k] + (1.0f - beta1) * gi;
      vv[k] = beta2 * vv[k] + (1.0f - beta2) * gi * gi;
      float w = pp[k] * it.decay - it.step_size * (mm[k] / (sqrtf(vv[k]) * it.inv_sqrt_bc2 + eps));
      if (it.clamp_max_enabled) w = fminf(w, it.clamp_max);
      pp[k] = w;
    }
    *(float4*)(p + base) = p4;
    *(float4*)(m + base) = m4;
    *(float4*)(v + base) = v4;
    return;
  }
  for (long i = base; i < base + 4 && i < it.n; ++i) {
    const float gi = g[i] * clip;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * it.inv_sqrt_bc2 + eps;
    float w = p[i] * it.decay;                    // decoupled weight decay applied BEFORE the Adam update (:168)
    w -= it.step_size * (mi / denom);
    if (it.clamp_max_enabled) w = fminf(w, it.clamp_max);
    p[i] = w;
  }
}

}  // namespace

extern "C" {

int sc_grad_sqnorm_multi(const sc_opt_item* items_dev, int n_items, int64_t total_blocks, float* out_sqnorm, void* stream) {
  SC_CHECK_ARG(items_dev && n_items > 0 && total_blocks > 0 && out_sqnorm, "sc_grad_sqnorm_multi: bad args");
  sc_count_launch(1);
  grad_sqnorm_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(items_dev, n_items, out_sqnorm);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_adamw_multi(const sc_opt_item* items_dev, int n_items, int64_t total_blocks, const float* sqnorm, float max_norm,
                   float beta1, float beta2, float eps, void* stream) {
  SC_CHECK_ARG(items_dev && n_items > 0 && total_blocks > 0 && sqnorm, "sc_adamw_multi: bad args");
  sc_count_launch(1);
  adamw_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(items_dev, n_items, sqnorm, max_norm, beta1, beta2, eps);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

}  // extern "C"
