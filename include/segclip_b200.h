/* libsegclip_b200 -- C ABI of the B200-native SegCLIP hot path (sm_100a).
 *
 * Drop-in boundary: the reference (ArrowLuo/SegCLIP) has no native code; its hot path is
 * `loss = model(input_ids, segment_ids, input_mask, image, image_seg=)` (main_task_align.py:312)
 * followed by `loss.backward()` (:321).  Every PyTorch library call below that surface is replaced
 * by one of the entry points declared here.  The reference-side binding is a ctypes stub
 * (INTEGRATION.md); segclip_b200/_lib.py is that stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory unless named host_*;
 *  - the caller (PyTorch) owns all memory, the library never allocates on the hot path and never
 *    synchronises the device; all work is ordered on the `stream` argument (a cudaStream_t);
 *  - every function returns 0 or a negative SC_ERR_* code; sc_last_error() gives the text
 *    (thread-local);
 *  - activations are row-major [rows, features]; `dtype` arguments are SC_F32 or SC_BF16.
 */
#ifndef SEGCLIP_B200_H
#define SEGCLIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_ERR_INVALID (-1)
#define SC_ERR_CUDA (-2)
#define SC_ERR_UNSUPPORTED (-3)

#define SC_F32 0
#define SC_BF16 1

#define SC_ACT_NONE 0
#define SC_ACT_QUICKGELU 1 /* x*sigmoid(1.702x): modules/module_clip_util.py:134-136 */
#define SC_ACT_GELU_ERF 2  /* nn.GELU():        modules/module_seg_vit.py:128, module_mae.py:151 */

#define SC_ABI_VERSION 1

const char* sc_last_error(void);
int sc_abi_version(void);
/* number of kernels launched by this library in this process so far (bench.py: gpu_launches) */
long long sc_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM with fused epilogue.  Replaces every nn.Linear / `@ proj` / F.conv2d(stride=kernel) /
 * grouped 1x1 Conv1d on the path and their autograd backward (dgrad, wgrad):
 *   modules/module_seg_vit.py:166-172,189 (in_proj/out_proj/c_fc/c_proj), :266-269 (k_conv/v_conv),
 *   modules/module_clip_vtransformer.py:21,56 (conv1), modules/module_clip.py:92,133 (proj),
 *   modules/module_mae.py:118-121,153-155,223,243 (decoder Linear layers).
 *
 *   acc[m,n] = sum_k A(m,k) * B(n,k)
 *   v        = alpha*acc + bias[n] + rowbias[ridx(m), n]
 *   if C2: C2[m,n] = v                       (pre-activation copy, saved for backward)
 *   v        = act(v) + residual[m,n]
 *   C[m,n]   = v   (or C[m,n] += v when accumulate; fp32 C only; atomic when split_k > 1)
 *
 * A(m,k) = A[m*lda + k] (trans_a = 0, "K-major") or A[k*lda + m] (trans_a = 1, "MN-major");
 * B(n,k) = B[n*ldb + k] (trans_b = 0)             or B[k*ldb + n] (trans_b = 1).
 * forward  Y = X W^T + b : A=X, B=W,  trans 0/0;   dgrad dX = dY W : A=dY, B=W, trans 0/1;
 * wgrad   dW = dY^T X    : A=dY, B=X, trans 1/1 (M=out_features, N=in_features, K=rows).
 *
 * in_dtype SC_BF16 -> tcgen05.mma kind::f16 (bf16 operands via TMA, fp32 accumulators in TMEM);
 * in_dtype SC_F32  -> exact fp32 FMA kernel (parity mode, tiny problems).
 * ridx(m) = rowbias_idx ? rowbias_idx[m] : m % rowbias_mod.
 */
typedef struct {
  int32_t M, N, K;
  int32_t in_dtype; /* dtype of A and B */
  int32_t trans_a, trans_b;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  float alpha;
  const float* bias;          /* [N] or NULL */
  const float* rowbias;       /* [*, N] or NULL, leading dimension ld_rowbias */
  int64_t ld_rowbias;
  const int32_t* rowbias_idx; /* [M] or NULL */
  int32_t rowbias_mod;
  int32_t act;
  const float* residual; /* fp32 [M, N] or NULL */
  int64_t ldr;
  void* C;
  int64_t ldc;
  int32_t c_dtype;
  void* C2; /* optional, same ldc */
  int32_t c2_dtype;
  int32_t accumulate;
  int32_t split_k;       /* 0/1 = none; >1 requires accumulate into fp32 C (atomic adds) */
  int32_t force_simt;    /* debugging / cross-checking: run the bf16 problem on the FMA kernel */
} sc_gemm_desc;

int sc_gemm(const sc_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGCLIP_B200_H */
