// Exact-fp32 FMA GEMM with the same fused epilogue as the tcgen05 kernel.  Used for the fp32 parity
// mode, for problems too small / misaligned for TMA, and to cross-check the tensor-core kernel.
#include "common.cuh"
#include "epilogue.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, long lda, int ta, const T* __restrict__ B,
                                                        long ldb, int tb, int K, EpiParams ep) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
    // 64x16 elements per operand, 256 threads -> 4 each; index so the contiguous dim is fastest
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int mm, kk;
      if (ta) { mm = idx & 63; kk = idx >> 6; } else { kk = idx & 15; mm = idx >> 4; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < ep.M && gk < K) v = to_f32(ta ? A[(long)gk * lda + gm] : A[(long)gm * lda + gk]);
      sA[kk][mm] = v;
      int nn;
      if (tb) { nn = idx & 63; kk = idx >> 6; } else { kk = idx & 15; nn = idx >> 4; }
      int gn = n0 + nn;
      gk = k0 + kk;
      v = 0.f;
      if (gn < ep.N && gk < K) v = to_f32(tb ? B[(long)gk * ldb + gn] : B[(long)gn * ldb + gk]);
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= ep.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < ep.N) epi_store_scalar(ep, m, n, acc[i][j]);
    }
  }
}

}  // namespace

extern void sc_count_kernel(int kind, int n);

int sc_gemm_simt(const sc_gemm_desc* d, cudaStream_t st) {
  sc_gemm_desc dd = *d;
  dd.split_k = 1;
  EpiParams ep = make_epi(&dd);
  dim3 grid(ceil_div(d->N, TN), ceil_div(d->M, TM));
  sc_count_kernel(SC_K_GEMM_SIMT, 1);
  if (d->in_dtype == SC_F32)
    gemm_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)d->A, d->lda, d->trans_a, (const float*)d->B, d->ldb,
                                                   d->trans_b, d->K, ep);
  else
    gemm_simt_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)d->A, d->lda, d->trans_a, (const bf16*)d->B, d->ldb,
                                                  d->trans_b, d->K, ep);
  SC_LAUNCH_CHECK();
  return SC_OK;
}
