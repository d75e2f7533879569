// Loss heads of SegCLIP.forward (modules/modeling.py:201-252): L2-normalised InfoNCE with rank-offset
// labels and global-LSE backward, superpixel-KL, masked-patch MSE of the MAE decoder, plus the MAE
// decoder's un-shuffle.  All fp32; each loss term is accumulated into one device scalar (no host sync).
#include "common.cuh"

extern void sc_count_launch(int n);

namespace {

constexpr int G = 8;

SC_DEVINL float block_sum(float v, float* red) {  // 256 threads
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}

// y = x / ||x||  (modules/modeling.py:341-345); one warp per row
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm, int rows, int E) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int d = lane; d < E; d += 32) { float v = x[(long)row * E + d]; s = fmaf(v, v, s); }
  const float inv = 1.0f / sqrtf(warp_sum(s));
  for (int d = lane; d < E; d += 32) y[(long)row * E + d] = x[(long)row * E + d] * inv;
  if (lane == 0) inv_norm[row] = inv;
}
// dx = (dy - y <y,dy>) * inv_norm
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv_norm,
                                  float* __restrict__ dx, int rows, int E) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int d = lane; d < E; d += 32) s = fmaf(y[(long)row * E + d], dy[(long)row * E + d], s);
  s = warp_sum(s);
  const float inv = inv_norm[row];
  for (int d = lane; d < E; d += 32) dx[(long)row * E + d] = (dy[(long)row * E + d] - y[(long)row * E + d] * s) * inv;
}

SC_DEVINL float logit_scale_value(const float* p) { return fminf(expf(*p), 100.0f); }  // modeling.py:350

// Per row i of raw = t_loc v_all^T (cosines): lse_i = logsumexp_j(s*raw_ij); loss += 0.5/B * (lse_i - s*raw_i,label)
__global__ void __launch_bounds__(256) ce_lse_kernel(const float* __restrict__ raw, int B, int N, int label_off,
                                                      const float* __restrict__ scale_param, float* __restrict__ lse,
                                                      float* __restrict__ loss) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float s = logit_scale_value(scale_param);
  const float* r = raw + (long)i * N;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += 256) mx = fmaxf(mx, s * r[j]);
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  float sum = 0.f;
  for (int j = threadIdx.x; j < N; j += 256) sum += expf(s * r[j] - mx);
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) {
    const float l = mx + logf(sum);
    lse[i] = l;
    atomicAdd(loss, (l - s * r[label_off + i]) * (0.5f / B));
  }
}

// In place: raw_ij <- s * [ (exp(s raw_ij - lse_own_i) - d_ij) + (exp(s raw_ij - lse_other_all_j) - d_ij) ] * (0.5/B)
// (own-row softmax gradient + the column term of the transposed direction); also
// dscale += sum_ij own-row-grad_ij * raw_ij * (dexp(p)/dp through the clamp).
__global__ void __launch_bounds__(256) ce_grad_kernel(float* __restrict__ raw, int B, int N, int label_off,
                                                       const float* __restrict__ scale_param, const float* __restrict__ lse_own,
                                                       const float* __restrict__ lse_other_all, float gscale,
                                                       float* __restrict__ dscale_param) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float p = *scale_param;
  const float s = fminf(expf(p), 100.0f);
  const float ds_dp = expf(p) <= 100.0f ? expf(p) : 0.f;
  float* r = raw + (long)i * N;
  const float lo = lse_own[i];
  const float c = gscale * 0.5f / B;
  float acc = 0.f;
  for (int j = threadIdx.x; j < N; j += 256) {
    const float x = r[j];
    const float onehot = (j == label_off + i) ? 1.f : 0.f;
    const float g1 = (expf(s * x - lo) - onehot) * c;
    const float g2 = (expf(s * x - lse_other_all[j]) - onehot) * c;
    acc = fmaf(g1, x, acc);
    r[j] = s * (g1 + g2);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(dscale_param, acc * ds_dp);
}

// ---------------------------------------------------------------- superpixel KL (modeling.py:212-224)
// one CTA per sample; hard_attn treated as the exact one-hot of idx (it differs by <= 1 ulp).
__global__ void __launch_bounds__(256) superpixel_kl_kernel(const int* __restrict__ idx, const long long* __restrict__ seg, int L,
                                                             float coef, float gscale, float* __restrict__ loss,
                                                             float* __restrict__ d_hard) {
  extern __shared__ float sm[];
  int* sidx = (int*)sm;                  // [L]
  long long* sseg = (long long*)(sidx + ((L + 1) & ~1));   // [L] (8-byte aligned)
  float* sgm = (float*)(sseg + L);       // [L][G]
  float* sga = sgm + L * G;              // [L][G]
  float* sn = sga + L * G;               // [L]
  __shared__ float red[8];
  const int b = blockIdx.x;
  for (int l = threadIdx.x; l < L; l += 256) {
    sidx[l] = idx[(long)b * L + l];
    sseg[l] = seg[(long)b * L + l];
  }
  __syncthreads();
  float f = 0.f;
  for (int l = threadIdx.x; l < L; l += 256) {
    float cnt[G];
#pragma unroll
    for (int c = 0; c < G; ++c) cnt[c] = 0.f;
    float n = 0.f;
    const long long me = sseg[l];
    for (int j = 0; j < L; ++j) {
      if (sseg[j] == me) {
        n += 1.f;
        const int cj = sidx[j];
#pragma unroll
        for (int c = 0; c < G; ++c) cnt[c] += (cj == c) ? 1.f : 0.f;
      }
    }
    sn[l] = n;
    // m = cnt / max(n,1); a = onehot(idx_l); p = softmax(m), q = softmax(a)
    float m[G], mxm = -INFINITY;
#pragma unroll
    for (int c = 0; c < G; ++c) { m[c] = cnt[c] / fmaxf(n, 1.f); mxm = fmaxf(mxm, m[c]); }
    float sp = 0.f;
#pragma unroll
    for (int c = 0; c < G; ++c) sp += expf(m[c] - mxm);
    const float lsp = mxm + logf(sp);
    const float lsq = 1.0f + logf(expf(0.f) + (G - 1) * expf(-1.f));  // logsumexp of a one-hot, max = 1
    const int me_c = sidx[l];
    float r[G], pp[G], qq[G], pr = 0.f, qr = 0.f;
#pragma unroll
    for (int c = 0; c < G; ++c) {
      const float lp = m[c] - lsp;
      const float lq = ((c == me_c) ? 1.f : 0.f) - lsq;
      pp[c] = expf(lp);
      qq[c] = expf(lq);
      r[c] = lp - lq;
      pr += pp[c] * r[c];
      qr += qq[c] * r[c];
    }
    f += pr - qr;   // sum_c (p-q)(lp-lq)
#pragma unroll
    for (int c = 0; c < G; ++c) {
      sgm[l * G + c] = pp[c] * (r[c] - pr) + pp[c] - qq[c];
      sga[l * G + c] = qq[c] * (-r[c] + qr) + qq[c] - pp[c];
    }
  }
  f = block_sum(f, red);
  if (threadIdx.x == 0) atomicAdd(loss, f * coef);
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += 256) {
    float acc[G];
#pragma unroll
    for (int c = 0; c < G; ++c) acc[c] = 0.f;
    const long long me = sseg[l];
    for (int j = 0; j < L; ++j)
      if (sseg[j] == me) {
#pragma unroll
        for (int c = 0; c < G; ++c) acc[c] += sgm[j * G + c];
      }
    const float inv_n = 1.0f / fmaxf(sn[l], 1.f);
#pragma unroll
    for (int c = 0; c < G; ++c) d_hard[((long)b * G + c) * L + l] = (sga[l * G + c] + acc[c] * inv_n) * coef * gscale;
  }
}

// ---------------------------------------------------------------- MAE decoder glue (module_mae.py:304-313)
template <typename T>
__global__ void __launch_bounds__(128) mae_unshuffle_kernel(const T* __restrict__ emb, const float* __restrict__ mask_token,
                                                             const int* __restrict__ ids_restore, const float* __restrict__ pos,
                                                             float* __restrict__ x, int L1, int keep, int D) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int r = ids_restore[(long)b * L1 + i];
  for (int d = threadIdx.x; d < D; d += 128) {
    const float v = (r < keep) ? to_f32(emb[((long)b * keep + r) * D + d]) : mask_token[d];
    x[((long)b * L1 + i) * D + d] = v + pos[(long)i * D + d];
  }
}
template <typename T>
__global__ void __launch_bounds__(128) mae_unshuffle_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ ids_restore,
                                                                 T* __restrict__ d_emb, float* __restrict__ d_mask_token, int L1,
                                                                 int keep, int D) {
  extern __shared__ int sr[];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < L1; i += 128) sr[i] = ids_restore[(long)b * L1 + i];
  __syncthreads();
  const int d = blockIdx.x * 128 + threadIdx.x;
  if (d >= D) return;
  float acc = 0.f;
  for (int i = 0; i < L1; ++i) {
    const float g = dx[((long)b * L1 + i) * D + d];
    const int r = sr[i];
    if (r < keep) d_emb[((long)b * keep + r) * D + d] = from_f32<T>(g);
    else acc += g;
  }
  atomicAdd(d_mask_token + d, acc);
}

// loss += sum_{masked (b,l)} mean_p (pred - target)^2 / (#masked); dpred = 2 (pred-target) mask / (P #masked)
// target = patchify(image) in (py, px, c) order (module_mae.py:18-29,322-328). one CTA per (patch, sample).
template <typename T>
__global__ void __launch_bounds__(256) mae_loss_kernel(const T* __restrict__ pred, const float* __restrict__ image,
                                                        const float* __restrict__ mask, int L1, int grid, int p, float inv_masked,
                                                        float gscale, float* __restrict__ loss, T* __restrict__ dpred) {
  __shared__ float red[8];
  const int l = blockIdx.x, b = blockIdx.y;   // l = token index including CLS
  const int P = 3 * p * p;
  const long row = ((long)b * L1 + l) * P;
  const float mk = (l == 0) ? 0.f : mask[(long)b * L1 + l];
  if (mk == 0.f) {   // block-uniform
    for (int k = threadIdx.x; k < P; k += 256) dpred[row + k] = from_f32<T>(0.f);
    return;
  }
  const int res = grid * p, gy = (l - 1) / grid, gx = (l - 1) % grid;
  float acc = 0.f;
  for (int k = threadIdx.x; k < P; k += 256) {
    const int c = k % 3, px = (k / 3) % p, py = k / (3 * p);
    const float t = image[(((long)b * 3 + c) * res + gy * p + py) * res + gx * p + px];
    const float df = to_f32(pred[row + k]) - t;
    acc = fmaf(df, df, acc);
    dpred[row + k] = from_f32<T>(2.f * df * inv_masked / P * gscale);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc / P * inv_masked);
}

}  // namespace

extern "C" {

int sc_l2norm_fwd(const float* x, float* y, float* inv_norm, int rows, int E, void* stream) {
  SC_CHECK_ARG(x && y && inv_norm, "sc_l2norm_fwd: null pointer");
  sc_count_launch(1);
  l2norm_fwd_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, y, inv_norm, rows, E);
  SC_LAUNCH_CHECK();
  return SC_OK;
}
int sc_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int rows, int E, void* stream) {
  SC_CHECK_ARG(dy && y && inv_norm && dx, "sc_l2norm_bwd: null pointer");
  sc_count_launch(1);
  l2norm_bwd_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(dy, y, inv_norm, dx, rows, E);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_ce_lse(const float* raw, int B, int N, int label_off, const float* logit_scale_param, float* lse, float* loss,
              void* stream) {
  SC_CHECK_ARG(raw && logit_scale_param && lse && loss && label_off + B <= N, "sc_ce_lse: bad args");
  sc_count_launch(1);
  ce_lse_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(raw, B, N, label_off, logit_scale_param, lse, loss);
  SC_LAUNCH_CHECK();
  return SC_OK;
}
int sc_ce_grad(float* raw, int B, int N, int label_off, const float* logit_scale_param, const float* lse_own,
               const float* lse_other_all, float gscale, float* d_logit_scale_param, void* stream) {
  SC_CHECK_ARG(raw && logit_scale_param && lse_own && lse_other_all && d_logit_scale_param, "sc_ce_grad: null pointer");
  sc_count_launch(1);
  ce_grad_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(raw, B, N, label_off, logit_scale_param, lse_own, lse_other_all,
                                                       gscale, d_logit_scale_param);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_superpixel_kl(const int32_t* idx, const int64_t* seg, int B, int L, float gscale, float* loss, float* d_hard,
                     void* stream) {
  SC_CHECK_ARG(idx && seg && loss && d_hard, "sc_superpixel_kl: null pointer");
  const size_t smem = sizeof(int) * ((L + 1) & ~1) + sizeof(long long) * L + sizeof(float) * (2 * L * G + L);
  const float coef = 0.5f / ((float)B * L * G);
  sc_count_launch(1);
  superpixel_kl_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(idx, (const long long*)seg, L, coef, gscale, loss, d_hard);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_mae_unshuffle(const void* emb, int emb_dtype, const float* mask_token, const int32_t* ids_restore, const float* pos,
                     float* x, int B, int L1, int keep, int D, void* stream) {
  SC_CHECK_ARG(emb && mask_token && ids_restore && pos && x, "sc_mae_unshuffle: null pointer");
  sc_count_launch(1);
  if (emb_dtype == SC_F32)
    mae_unshuffle_kernel<float><<<dim3(L1, B), 128, 0, (cudaStream_t)stream>>>((const float*)emb, mask_token, ids_restore, pos, x, L1, keep, D);
  else
    mae_unshuffle_kernel<bf16><<<dim3(L1, B), 128, 0, (cudaStream_t)stream>>>((const bf16*)emb, mask_token, ids_restore, pos, x, L1, keep, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}
int sc_mae_unshuffle_bwd(const float* dx, const int32_t* ids_restore, void* d_emb, int emb_dtype, float* d_mask_token, int B,
                         int L1, int keep, int D, void* stream) {
  SC_CHECK_ARG(dx && ids_restore && d_emb && d_mask_token, "sc_mae_unshuffle_bwd: null pointer");
  sc_count_launch(1);
  dim3 grid(ceil_div(D, 128), B);
  if (emb_dtype == SC_F32)
    mae_unshuffle_bwd_kernel<float><<<grid, 128, L1 * sizeof(int), (cudaStream_t)stream>>>(dx, ids_restore, (float*)d_emb, d_mask_token, L1, keep, D);
  else
    mae_unshuffle_bwd_kernel<bf16><<<grid, 128, L1 * sizeof(int), (cudaStream_t)stream>>>(dx, ids_restore, (bf16*)d_emb, d_mask_token, L1, keep, D);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

int sc_mae_loss(const void* pred, int dtype, const float* image, const float* mask, int B, int L1, int keep, int grid,
                int patch, float gscale, float* loss, void* dpred, void* stream) {
  SC_CHECK_ARG(pred && image && mask && loss && dpred, "sc_mae_loss: null pointer");
  SC_CHECK_ARG(L1 == grid * grid + 1 && keep < L1, "sc_mae_loss: bad shape");
  const float inv_masked = 1.0f / ((float)B * (L1 - keep));
  sc_count_launch(1);
  if (dtype == SC_F32)
    mae_loss_kernel<float><<<dim3(L1, B), 256, 0, (cudaStream_t)stream>>>((const float*)pred, image, mask, L1, grid, patch, inv_masked, gscale, loss, (float*)dpred);
  else
    mae_loss_kernel<bf16><<<dim3(L1, B), 256, 0, (cudaStream_t)stream>>>((const bf16*)pred, image, mask, L1, grid, patch, inv_masked, gscale, loss, (bf16*)dpred);
  SC_LAUNCH_CHECK();
  return SC_OK;
}

}  // extern "C"
