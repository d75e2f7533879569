// Generic softmax-attention core, forward and backward, fp32 math on CUDA cores.
// Replaces the bmm / softmax / bmm inside nn.MultiheadAttention (modules/module_seg_vit.py:189,215,
// modules/module_clip_ttransformer.py:46) and the timm Attention of the MAE decoder
// (modules/module_mae.py:122-135), for every shape on the path: self-attention L in {196,48,8,77,197},
// head dim 64 / 48, causal text mask (module_clip_util.py:199-205), and the centre cross-attention with
// either K/V layout -- the layout is just a pair of strides here.
//
// One CTA per (head, batch slot); K and V (fwd, dQ pass) or Q and dO (dK/dV pass) staged once in
// shared memory (padded rows: conflict-free), one warp per query (resp. key) row.
// Addressing: element (b, i, h, d) of X lives at X + b*bs + i*rs + h*hd + d.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;

template <typename T>
SC_DEVINL void stage_rows(float* dst, int ldd, const T* src, long rs, int rows, int hd) {
  // dst[r*ldd + d] = src[r*rs + d]
  for (int idx = threadIdx.x; idx < rows * hd; idx += ATT_THREADS) {
    int r = idx / hd, d = idx - r * hd;
    dst[r * ldd + d] = to_f32(src[(long)r * rs + d]);
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(sc_attn_desc a) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = a.hd, ldk = hd + 1, Lk = a.Lk, Lq = a.Lq;
  float* sK = sm;
  float* sV = sK + Lk * ldk;
  float* sP = sV + Lk * ldk;            // [ATT_WARPS][Lk]
  float* sQ = sP + ATT_WARPS * Lk;      // [ATT_WARPS][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows<T>(sK, ldk, (const T*)a.k + (long)b * a.k_bs + h * hd, a.k_rs, Lk, hd);
  stage_rows<T>(sV, ldk, (const T*)a.v + (long)b * a.v_bs + h * hd, a.v_rs, Lk, hd);
  __syncthreads();
  float* p = sP + warp * Lk;
  float* qs = sQ + warp * 64;
  for (int i = warp; i < Lq; i += ATT_WARPS) {
    const T* q = (const T*)a.q + (long)b * a.q_bs + (long)i * a.q_rs + h * hd;
    for (int d = lane; d < hd; d += 32) qs[d] = to_f32(q[d]) * a.scale;
    __syncwarp();
    const int kmax = a.causal ? min(Lk, i + 1) : Lk;
    float mx = -INFINITY;
    for (int j = lane; j < kmax; j += 32) {
      const float* kr = sK + j * ldk;
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < hd; ++d) s = fmaf(qs[d], kr[d], s);
      p[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < kmax; j += 32) {
      float e = __expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    float o0 = 0.f, o1 = 0.f;
    const bool has1 = lane + 32 < hd;
    for (int j = 0; j < kmax; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, sV[j * ldk + lane], o0);
      if (has1) o1 = fmaf(pj, sV[j * ldk + lane + 32], o1);
    }
    T* o = (T*)a.o + (long)b * a.o_bs + (long)i * a.o_rs + h * hd;
    if (lane < hd) o[lane] = from_f32<T>(o0 * inv);
    if (has1) o[lane + 32] = from_f32<T>(o1 * inv);
    if (lane == 0 && a.lse) a.lse[((long)b * a.H + h) * Lq + i] = mx + __logf(sum);
    __syncwarp();
  }
}

// dQ pass: one warp per query row; K, V in smem.
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(sc_attn_bwd_desc g) {
  const sc_attn_desc& a = g.fwd;
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = a.hd, ldk = hd + 1, Lk = a.Lk, Lq = a.Lq;
  float* sK = sm;
  float* sV = sK + Lk * ldk;
  float* sP = sV + Lk * ldk;
  float* sQ = sP + ATT_WARPS * Lk;       // [W][64] scaled q
  float* sdO = sQ + ATT_WARPS * 64;      // [W][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows<T>(sK, ldk, (const T*)a.k + (long)b * a.k_bs + h * hd, a.k_rs, Lk, hd);
  stage_rows<T>(sV, ldk, (const T*)a.v + (long)b * a.v_bs + h * hd, a.v_rs, Lk, hd);
  __syncthreads();
  float* p = sP + warp * Lk;
  float* qs = sQ + warp * 64;
  float* dos = sdO + warp * 64;
  for (int i = warp; i < Lq; i += ATT_WARPS) {
    const long qoff = (long)b * a.q_bs + (long)i * a.q_rs + h * hd;
    const long ooff = (long)b * a.o_bs + (long)i * a.o_rs + h * hd;
    float dpart = 0.f;
    for (int d = lane; d < hd; d += 32) {
      qs[d] = to_f32(((const T*)a.q)[qoff + d]) * a.scale;
      float dv = to_f32(((const T*)g.d_o)[ooff + d]);
      dos[d] = dv;
      dpart += dv * to_f32(((const T*)a.o)[ooff + d]);
    }
    const float delta = warp_sum(dpart);
    const float lse = a.lse[((long)b * a.H + h) * Lq + i];
    __syncwarp();
    const int kmax = a.causal ? min(Lk, i + 1) : Lk;
    for (int j = lane; j < kmax; j += 32) {
      const float* kr = sK + j * ldk;
      const float* vr = sV + j * ldk;
      float s = 0.f, dp = 0.f;
#pragma unroll 8
      for (int d = 0; d < hd; ++d) {
        s = fmaf(qs[d], kr[d], s);
        dp = fmaf(dos[d], vr[d], dp);
      }
      p[j] = __expf(s - lse) * (dp - delta);   // dS_ij
    }
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    const bool has1 = lane + 32 < hd;
    for (int j = 0; j < kmax; ++j) {
      const float ds = p[j];
      a0 = fmaf(ds, sK[j * ldk + lane], a0);
      if (has1) a1 = fmaf(ds, sK[j * ldk + lane + 32], a1);
    }
    T* dq = (T*)g.d_q + qoff;
    if (lane < hd) dq[lane] = from_f32<T>(a0 * a.scale);
    if (has1) dq[lane + 32] = from_f32<T>(a1 * a.scale);
    __syncwarp();
  }
}

// dK/dV pass: one warp per key row; Q (scaled), dO in smem, delta/lse per query in smem.
template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_kernel(sc_attn_bwd_desc g) {
  const sc_attn_desc& a = g.fwd;
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = a.hd, ldq = hd + 1, Lk = a.Lk, Lq = a.Lq;
  float* sQ = sm;                        // [Lq][ldq] scaled
  float* sdO = sQ + Lq * ldq;            // [Lq][ldq]
  float* sDelta = sdO + Lq * ldq;        // [Lq]
  float* sLse = sDelta + Lq;             // [Lq]
  float* sP = sLse + Lq;                 // [W][Lq]
  float* sdS = sP + ATT_WARPS * Lq;      // [W][Lq]
  float* sKr = sdS + ATT_WARPS * Lq;     // [W][64]
  float* sVr = sKr + ATT_WARPS * 64;     // [W][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows<T>(sQ, ldq, (const T*)a.q + (long)b * a.q_bs + h * hd, a.q_rs, Lq, hd);
  stage_rows<T>(sdO, ldq, (const T*)g.d_o + (long)b * a.o_bs + h * hd, a.o_rs, Lq, hd);
  __syncthreads();
  for (int i = warp; i < Lq; i += ATT_WARPS) {
    const T* o = (const T*)a.o + (long)b * a.o_bs + (long)i * a.o_rs + h * hd;
    float dpart = 0.f;
    for (int d = lane; d < hd; d += 32) dpart += sdO[i * ldq + d] * to_f32(o[d]);
    dpart = warp_sum(dpart);
    if (lane == 0) {
      sDelta[i] = dpart;
      sLse[i] = a.lse[((long)b * a.H + h) * Lq + i];
    }
  }
  __syncthreads();
  float* p = sP + warp * Lq;
  float* ds = sdS + warp * Lq;
  float* kr = sKr + warp * 64;
  float* vr = sVr + warp * 64;
  for (int j = warp; j < Lk; j += ATT_WARPS) {
    const long koff = (long)b * a.k_bs + (long)j * a.k_rs + h * hd;
    const long voff = (long)b * a.v_bs + (long)j * a.v_rs + h * hd;
    for (int d = lane; d < hd; d += 32) {
      kr[d] = to_f32(((const T*)a.k)[koff + d]);
      vr[d] = to_f32(((const T*)a.v)[voff + d]);
    }
    __syncwarp();
    const int imin = a.causal ? j : 0;
    for (int i = imin + lane; i < Lq; i += 32) {
      const float* qr = sQ + i * ldq;
      const float* dor = sdO + i * ldq;
      float s = 0.f, dp = 0.f;
#pragma unroll 8
      for (int d = 0; d < hd; ++d) {
        s = fmaf(qr[d] * a.scale, kr[d], s);
        dp = fmaf(dor[d], vr[d], dp);
      }
      const float pij = __expf(s - sLse[i]);
      p[i] = pij;
      ds[i] = pij * (dp - sDelta[i]);
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    const bool has1 = lane + 32 < hd;
    for (int i = imin; i < Lq; ++i) {
      const float pi = p[i], dsi = ds[i];
      v0 = fmaf(pi, sdO[i * ldq + lane], v0);
      k0 = fmaf(dsi, sQ[i * ldq + lane], k0);
      if (has1) {
        v1 = fmaf(pi, sdO[i * ldq + lane + 32], v1);
        k1 = fmaf(dsi, sQ[i * ldq + lane + 32], k1);
      }
    }
    T* dk = (T*)g.d_k + koff;
    T* dv = (T*)g.d_v + voff;
    if (lane < hd) {
      dk[lane] = from_f32<T>(k0 * a.scale);
      dv[lane] = from_f32<T>(v0);
    }
    if (has1) {
      dk[lane + 32] = from_f32<T>(k1 * a.scale);
      dv[lane + 32] = from_f32<T>(v1);
    }
    __syncwarp();
  }
}

int check_desc(const sc_attn_desc* a, const char* who) {
  SC_CHECK_ARG(a->q && a->k && a->v && a->o, "%s: null pointer", who);
  SC_CHECK_ARG(a->hd > 0 && a->hd <= 64, "%s: head dim %d unsupported (<= 64)", who, a->hd);
  SC_CHECK_ARG(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "%s: bad shape", who);
  SC_CHECK_ARG(a->dtype == SC_F32 || a->dtype == SC_BF16, "%s: bad dtype", who);
  SC_CHECK_ARG(a->B <= 65535, "%s: batch %d > 65535", who, a->B);
  return SC_OK;
}

template <typename K>
int set_smem(K kern, size_t bytes, const char* who) {
  if (bytes > 227 * 1024) {
    sc_set_error("%s: needs %zu bytes of shared memory (> 227 KB); sequence too long for this kernel", who, bytes);
    return SC_ERR_UNSUPPORTED;
  }
  SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SC_OK;
}

}  // namespace

extern void sc_count_launch(int n);
extern void sc_count_kernel(int kind, int n);
bool sc_attn_mma_supported(const sc_attn_desc* a);
int sc_attention_fwd_mma(const sc_attn_desc* a, cudaStream_t st);
int sc_attention_bwd_mma(const sc_attn_bwd_desc* g, float* delta, cudaStream_t st);
bool sc_attn_tc_supported(const sc_attn_desc* a);
int sc_attention_bwd_tc(const sc_attn_bwd_desc* g, float* delta, cudaStream_t st);
int sc_attention_fwd_tc(const sc_attn_desc* a, cudaStream_t st);

extern "C" int sc_attention_fwd(const sc_attn_desc* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int rc = check_desc(a, "sc_attention_fwd");
  if (rc) return rc;
  {
    static const int use_tc = [] { const char* e = getenv("SC_ATT_FWD_TC"); return e ? atoi(e) : 1; }();
    if (use_tc && !a->force_generic && a->lse && sc_attn_tc_supported(a) && a->B <= 65535 && a->H <= 65535)
      return sc_attention_fwd_tc(a, st);
  }
  if (!a->force_generic && a->lse && sc_attn_mma_supported(a)) return sc_attention_fwd_mma(a, st);
  const size_t smem = sizeof(float) * ((size_t)2 * a->Lk * (a->hd + 1) + (size_t)ATT_WARPS * a->Lk + ATT_WARPS * 64);
  dim3 grid(a->H, a->B);
  sc_count_kernel(SC_K_ATTN_GENERIC, 1);
  if (a->dtype == SC_F32) {
    if ((rc = set_smem(attn_fwd_kernel<float>, smem, "sc_attention_fwd"))) return rc;
    attn_fwd_kernel<float><<<grid, ATT_THREADS, smem, st>>>(*a);
  } else {
    if ((rc = set_smem(attn_fwd_kernel<bf16>, smem, "sc_attention_fwd"))) return rc;
    attn_fwd_kernel<bf16><<<grid, ATT_THREADS, smem, st>>>(*a);
  }
  SC_LAUNCH_CHECK();
  return SC_OK;
}

extern "C" int sc_colsum(const void* x, int dtype, int64_t ld, int64_t rows, int cols, float* out, void* stream);

// bias-gradient column sums of d_q / d_k / d_v for the paths whose kernels do not produce them (sample-major packed rows)
static int colsum_after(const sc_attn_bwd_desc* g, cudaStream_t st) {
  const sc_attn_desc* a = &g->fwd;
  if (!g->dq_colsum) return SC_OK;
  const int W = a->H * a->hd;
  SC_CHECK_ARG(a->q_bs == (long)a->Lq * a->q_rs && a->k_bs == (long)a->Lk * a->k_rs && a->v_bs == (long)a->Lk * a->v_rs,
               "sc_attention_bwd: fused bias-gradient column sums need sample-major packed q / k / v");
  int rc;
  if ((rc = sc_colsum(g->d_q, a->dtype, a->q_rs, (int64_t)a->B * a->Lq, W, g->dq_colsum, st))) return rc;
  if ((rc = sc_colsum(g->d_k, a->dtype, a->k_rs, (int64_t)a->B * a->Lk, W, g->dk_colsum, st))) return rc;
  return sc_colsum(g->d_v, a->dtype, a->v_rs, (int64_t)a->B * a->Lk, W, g->dv_colsum, st);
}

extern "C" int sc_attention_bwd(const sc_attn_bwd_desc* g, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const sc_attn_desc* a = &g->fwd;
  int rc = check_desc(a, "sc_attention_bwd");
  if (rc) return rc;
  SC_CHECK_ARG(g->d_o && g->d_q && g->d_k && g->d_v && a->lse, "sc_attention_bwd: null pointer");
  SC_CHECK_ARG((g->dq_colsum != nullptr) == (g->dk_colsum != nullptr) && (g->dq_colsum != nullptr) == (g->dv_colsum != nullptr),
               "sc_attention_bwd: dq/dk/dv_colsum come together");
  {
    static const int use_tc = [] { const char* e = getenv("SC_ATT_TC"); return e ? atoi(e) : 1; }();
    if (use_tc && !a->force_generic && g->delta_ws && sc_attn_tc_supported(a)) return sc_attention_bwd_tc(g, g->delta_ws, st);
  }
  if (!a->force_generic && g->delta_ws && sc_attn_mma_supported(a)) {
    if ((rc = sc_attention_bwd_mma(g, g->delta_ws, st))) return rc;
    return colsum_after(g, st);
  }
  const size_t smem_q = sizeof(float) * ((size_t)2 * a->Lk * (a->hd + 1) + (size_t)ATT_WARPS * a->Lk + 2 * ATT_WARPS * 64);
  const size_t smem_kv = sizeof(float) * ((size_t)2 * a->Lq * (a->hd + 1) + 2 * (size_t)a->Lq +
                                          2 * (size_t)ATT_WARPS * a->Lq + 2 * ATT_WARPS * 64);
  dim3 grid(a->H, a->B);
  sc_count_kernel(SC_K_ATTN_GENERIC, 2);
  if (a->dtype == SC_F32) {
    if ((rc = set_smem(attn_bwd_dq_kernel<float>, smem_q, "sc_attention_bwd"))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<float>, smem_kv, "sc_attention_bwd"))) return rc;
    attn_bwd_dq_kernel<float><<<grid, ATT_THREADS, smem_q, st>>>(*g);
    attn_bwd_dkv_kernel<float><<<grid, ATT_THREADS, smem_kv, st>>>(*g);
  } else {
    if ((rc = set_smem(attn_bwd_dq_kernel<bf16>, smem_q, "sc_attention_bwd"))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<bf16>, smem_kv, "sc_attention_bwd"))) return rc;
    attn_bwd_dq_kernel<bf16><<<grid, ATT_THREADS, smem_q, st>>>(*g);
    attn_bwd_dkv_kernel<bf16><<<grid, ATT_THREADS, smem_kv, st>>>(*g);
  }
  SC_LAUNCH_CHECK();
  return colsum_after(g, st);
}
