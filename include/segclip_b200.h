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
#define SC_ACT_DERIV 3     /* mul_aux_act only: mul_aux already holds act'(pre-activation), see c2_is_act_grad */

#define SC_ABI_VERSION 1

const char* sc_last_error(void);
int sc_abi_version(void);
/* number of kernels launched by this library in this process so far (bench.py: gpu_launches) */
long long sc_launch_count(void);
/* per-kernel-family launch counters (tests assert that the benchmark's tensor-core kernels, not a fallback, ran) */
#define SC_K_GEMM_TC2 0     /* gemm_tc2_kernel: tcgen05 cta_group::2, 256 x 256 pair tiles */
#define SC_K_GEMM_TC1 1     /* gemm_tc_kernel: tcgen05 cta_group::1 */
#define SC_K_GEMM_SIMT 2    /* exact fp32 FMA kernel */
#define SC_K_ATTN_FWD_TC 3  /* tcgen05 attention forward */
#define SC_K_ATTN_BWD_TC 4  /* tcgen05 attention backward */
#define SC_K_ATTN_MMA 5     /* mma.sync attention (legacy tensor path) */
#define SC_K_ATTN_GENERIC 6 /* exact fp32 attention */
#define SC_K_GEMM_TC2_TAIL 7 /* gemm_tc2_kernel launches whose incomplete last wave ran as column slices (counted in addition to kind 0) */
#define SC_K_COUNT 8
long long sc_kernel_launches(int kind);

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
 *   v        = act(v) * act'(mul_aux[m,n]) + residual[m,n]     (act' term only if mul_aux)
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
  const void* residual; /* [M, N] or NULL, dtype residual_dtype (fp32, or bf16 with a bf16 C on the tensor-core path) */
  int64_t ldr;
  void* C;
  int64_t ldc;
  int32_t c_dtype;
  void* C2; /* optional, same ldc */
  int32_t c2_dtype;
  int32_t accumulate;
  int32_t split_k;       /* 0/1 = none; >1 requires accumulate into fp32 C (atomic adds) */
  int32_t force_simt;    /* debugging / cross-checking: run the bf16 problem on the FMA kernel */
  const void* mul_aux;   /* optional [M,N] (leading dim ldc): v *= act'(mul_aux) after act -- fused QuickGELU/GELU backward */
  int32_t mul_aux_dtype;
  int32_t mul_aux_act;
  float* colsum_out;     /* optional fp32 [N]: colsum_out[n] += sum_m C[m,n] (bias gradient fused into the dgrad that produces dY);
                            only with bf16 C on the tensor-core path */
  int32_t residual_dtype; /* SC_F32 (default) or SC_BF16 */
  /* optional per-head row dots of the result with a second matrix (bf16 C on the tensor-core path, head dim 64):
   *   dot_out[((m / dot_L) * (N / 64) + n / 64) * dot_L + m % dot_L] = sum over the 64 columns of head n / 64 of C[m,:] * dot_aux[m,:]
   * i.e. delta = rowsum(dO o O) of the attention backward (modules/module_seg_vit.py:189), produced by the out_proj dgrad
   * that computes dO instead of a separate pass over dO and O.  dot_aux: bf16 [M, N], leading dimension ldc. */
  const void* dot_aux;
  float* dot_out;
  int32_t dot_L;
  /* forward of an activated Linear whose backward is fused into the next dgrad: C2 receives act'(v) instead of the
   * pre-activation v (the epilogue has the sigmoid / erf at hand; the backward then multiplies with mul_aux_act =
   * SC_ACT_DERIV instead of re-evaluating the derivative for every element) */
  int32_t c2_is_act_grad;
} sc_gemm_desc;

int sc_gemm(const sc_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (fp32 statistics).  Replaces LayerNorm.forward modules/module_clip_util.py:126-132 and
 * the nn.LayerNorm calls at modules/module_seg_vit.py:288,297,300,312, modules/module_clip.py:86,129,
 * modules/module_mae.py:196-197,317 and their autograd backward.
 * y[map(r), :] = (x[r,:] - mean_r) * rstd_r * gamma + beta, with the optional row map
 *   map(r) = (r / in_group) * out_group + out_off + r % in_group   (in_group == 0: identity)
 * which writes straight into a concatenated buffer (kv_ = cat([q_feat, x]) module_seg_vit.py:294).
 */
typedef struct {
  int64_t rows;
  int32_t D;
  const void* x;
  int32_t x_dtype;
  const float* gamma;
  const float* beta;
  float eps;
  void* y;
  int32_t y_dtype;
  int32_t in_group, out_group, out_off;
  float* mean; /* [rows] saved statistics (may be NULL) */
  float* rstd;
} sc_ln_desc;
int sc_layernorm_fwd(const sc_ln_desc* d, void* stream);

/* dx[r,:] (+)= LN-backward(dy[map(r),:]);  dgamma/dbeta (fp32 [D], caller-zeroed) accumulate
 * atomically over rows and over calls. */
typedef struct {
  int64_t rows;
  int32_t D;
  const void* dy;
  int32_t dy_dtype;
  const void* x;
  int32_t x_dtype;
  const float* mean;
  const float* rstd;
  const float* gamma;
  void* dx; /* may be NULL (only parameter gradients wanted) */
  int32_t dx_dtype;
  int32_t accumulate_dx;
  void* dx_copy_bf16; /* optional bf16 copy of the final dx (operand of the next dgrad/wgrad GEMM) */
  float* dgamma;
  float* dbeta;
  int32_t in_group, out_group, out_off;
  float* dx_colsum; /* optional fp32 [D]: += column sums of the final dx (bias gradient of the Linear that produced x's
                       residual branch, fused here because this kernel is the producer of that dY) */
} sc_ln_bwd_desc;
int sc_layernorm_bwd(const sc_ln_bwd_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention core  o = softmax(scale * q k^T [+ causal mask]) v  per (batch slot, head).
 * Replaces the bmm/softmax/bmm inside nn.MultiheadAttention (modules/module_seg_vit.py:189,215;
 * modules/module_clip_ttransformer.py:46) and timm Attention (modules/module_mae.py:122-135).
 * Element (b, i, h, d) of X is X[b*X_bs + i*X_rs + h*hd + d] (strides in elements), so packed QKV
 * buffers and both centre-cross-attention K/V layouts ("torch18_flat" k_bs = row,
 * k_rs = B rows; "per_sample" k_bs = Lk rows, k_rs = row) are expressed without copies.
 * lse[b,h,i] = log sum_j exp(s_ij) is saved for the backward pass.
 * bf16 problems with head dim 32/48/64 and 16 <= L <= 1024 run flash-style on the tensor cores
 * (mma.sync m16n8k16, no [L,L] matrix in HBM); everything else (fp32 parity mode, the 8-query centre
 * cross-attention, tiny L) runs on an exact fp32 kernel.
 */
typedef struct {
  int32_t B, H, Lq, Lk, hd;
  int32_t dtype; /* of q, k, v, o and their gradients */
  int32_t causal;
  float scale;
  const void* q;
  int64_t q_bs, q_rs;
  const void* k;
  int64_t k_bs, k_rs;
  const void* v;
  int64_t v_bs, v_rs;
  void* o;
  int64_t o_bs, o_rs;
  float* lse; /* [B, H, Lq] */
  int32_t force_generic; /* debugging: skip the tensor-core kernels */
} sc_attn_desc;
int sc_attention_fwd(const sc_attn_desc* a, void* stream);

typedef struct {
  sc_attn_desc fwd; /* same pointers as the forward call (o and lse are inputs here) */
  const void* d_o;  /* strides of o */
  void* d_q;        /* strides of q, k, v respectively */
  void* d_k;
  void* d_v;
  float* delta_ws; /* fp32 scratch [B, H, Lq] (rowsum(dO o O)); required by the tensor-core kernels */
  /* optional fp32 [H*hd] each (all three or none): += column sums over all rows of d_q / d_k / d_v -- the bias gradient of
   * the projection that produced q, k, v (in_proj_bias, modules/module_seg_vit.py:189), fused into the kernel that writes
   * them (tcgen05 path: from the staged output tiles; other paths: a column-sum launch after the kernels) */
  float* dq_colsum;
  float* dk_colsum;
  float* dv_colsum;
  int32_t delta_ready; /* delta_ws already holds rowsum(dO o O) (written by sc_gemm's dot_out): skip the delta pass */
} sc_attn_bwd_desc;
int sc_attention_bwd(const sc_attn_bwd_desc* g, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise / data-movement glue (all HBM-bound, one pass).
 */
/* dx = dy * act'(pre): autograd backward of QuickGELU / nn.GELU (module_clip_util.py:134-136). */
int sc_act_bwd(const void* dy, int dy_dtype, const void* pre, int pre_dtype, void* dx, int dx_dtype, int64_t n, int act,
               void* stream);
/* dst = convert(src) * (scale_dev ? *scale_dev : 1): dtype conversion of residual-stream gradients and the final
 * `grad * grad_output` hand-over to autograd (scale read on the device, no host sync). */
int sc_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, const float* scale_dev, void* stream);
/* out[c] += sum_r x[r*ld + c] (atomic): bias gradients of every nn.Linear on the path. */
int sc_colsum(const void* x, int dtype, int64_t ld, int64_t rows, int cols, float* out, void* stream);

/* fp32 master parameters -> compute-dtype shadow copies, one launch for the whole model
 * (replaces the per-parameter `.type(self.dtype)` casts, modules/module_clip.py:84,109-112). */
typedef struct {
  const void* src; /* fp32 */
  void* dst;
  int64_t n;
  int64_t first_block; /* prefix sum of ceil(n / 1024) */
} sc_cast_item;
int sc_cast_multi(const sc_cast_item* items_dev, int n_items, int64_t total_blocks, int dst_dtype, void* stream);

/* nn.Conv1d(C, C, 1, groups) weight [C, C/groups] <-> dense block-diagonal [C, C]
 * (k_conv / v_conv, modules/module_seg_vit.py:266-269,299,302). reduce: dw += diag blocks of dense. */
int sc_blockdiag_expand(const float* w, void* dense, int C, int groups, int dtype, void* stream);
int sc_blockdiag_reduce(const float* dense_grad, float* dw, int C, int groups, void* stream);

/* Non-overlapping patch extraction for conv1 (kernel == stride, module_clip_vtransformer.py:21,56):
 * out[r*ld + c*p*p + y*p + x] = image[r / rows_per_img, c, (pid / grid_w)*p + y, (pid % grid_w)*p + x],
 * pid = patch_idx ? patch_idx[r] : r % rows_per_img (MAE pass: only kept patches are embedded); the image is
 * [*, 3, grid_h*p, grid_w*p] (training: square; inference also takes the 2x / rectangular sizes of the reference).
 * ld >= 3*p*p is the row pitch of `out` (padded to a multiple of 8 for TMA when 3*p*p is not, e.g. patch 14). */
int sc_im2col(const float* image, void* out, int out_dtype, int64_t ld, const int32_t* patch_idx, int64_t rows,
              int rows_per_img, int grid_h, int grid_w, int patch, void* stream);
/* Inference at another input size: bicubic (A = -0.75, align_corners = False) interpolation of the patch part of the
 * positional table [src_h, src_w, D] -> [dst_h, dst_w, D], fp32 (VisualTransformer.get_pos_embed,
 * modules/module_clip_vtransformer.py:35-53; same arithmetic as F.interpolate(mode='bicubic')). */
int sc_bicubic_resize(const float* src, float* dst, int src_h, int src_w, int dst_h, int dst_w, int D, void* stream);

/* "Next" row: uint8 image boundary.  out = (img/255 - mean[c]) / std[c] for a [*, 3, H, W] uint8
 * batch already on the device (hw = H*W; mean3/std3 are HOST arrays of 3 floats): the normalisation of
 * dataloaders/rawimage_util.py moved behind the H2D copy, so 1 byte per pixel crosses PCIe instead of 4 (or 8). */
int sc_u8_normalize(const uint8_t* img, float* out, int64_t n, int hw, const float* mean3, const float* std3, void* stream);

/* token_embedding(ids) + positional_embedding (modules/module_clip.py:109-112) and the flat row index
 * b*T + argmax_t ids[b,t] of the EOT token (:136). */
int sc_text_embed(const int64_t* ids, const float* tok, const float* pos, float* out, int32_t* eot_rows, int B, int T,
                  int W, void* stream);
/* out[r,:] = src[idx[r],:]; src in src_dtype (fp32, or the bf16 residual stream), out fp32 */
int sc_gather_rows(const void* src, int src_dtype, const int32_t* idx, float* out, int64_t rows, int D, void* stream);
int sc_scatter_rows(const float* src, const int32_t* idx, float* out, int64_t rows, int D, void* stream);
/* out[idx[r] + idx_offset, :] += src[r, :] (fp32 atomics).  Backward of an embedding lookup: nn.Embedding of the token ids
 * (modules/module_clip.py:107, idx int64) and the positional rows of the masked visual pass (module_clip_vtransformer.py:66-71,
 * idx int32 patch index, idx_offset 1 for the CLS row) when the reference's frozen stem is trained. */
int sc_scatter_add_rows(const void* src, int src_dtype, const void* idx, int idx_is_int64, int64_t idx_offset, float* out,
                        int64_t rows, int D, void* stream);

/* random_masking(keep_cls=True) (modules/module_clip_util.py:91-124) from an explicit uniform draw u
 * [B, L1]: ids_restore = argsort(argsort(noise)), ids_keep [B, keep], mask (1 = removed) and
 * patch_idx [B, keep-1] = ids_keep[:,1:] - 1. */
int sc_mae_mask(const float* u, int B, int L1, int keep, int32_t* ids_restore, int32_t* ids_keep, float* mask,
                int32_t* patch_idx, void* stream);

/* CLS := max over the G centre tokens (modules/module_seg_vit.py:441) and its backward routing. */
int sc_pool_max(const float* x, float* out, int32_t* arg, int B, int G, int D, void* stream);
int sc_pool_max_bwd(const float* dout, const int32_t* arg, float* dx, int B, int G, int D, void* stream);
/* cat([mean(x, 1), x], 1) (modules/modeling.py:244-245) and backward. */
int sc_mean_cat(const float* x, void* out, int out_dtype, int B, int n, int D, void* stream);
int sc_mean_cat_bwd(const void* dout, int dout_dtype, float* dx, int B, int n, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Patch -> centre aggregation (SemanticLearnerModule.forward, modules/module_seg_vit.py:304-312).
 *   logits[b,g,l] = <qf[b,g,:], k[b,l,:]>                                   (:304)
 *   y_soft = softmax((logits + Gumbel(u)) / tau, over g);  idx = argmax_g   (:221-237, tau = 0.9)
 *   soft   = softmax(logits, over g)                                        (:306)
 *   count[b,g] += #{l: idx = g}   (caller zeroes count)
 * forced_idx (may be NULL) teacher-forces the arg-max (bf16 gradient parity tests).
 */
typedef struct {
  int32_t B, G, L, D;
  const float* qf; /* [B,G,D] cross_ln output */
  const void* k;   /* [B,L,D] k_ln output */
  int32_t k_dtype;
  const float* u; /* [B,G,L] torch.rand draw behind the Gumbel noise; NULL = inference (plain softmax / arg-max of the logits) */
  float tau;
  const int32_t* forced_idx; /* [B,L] or NULL */
  float* logits;             /* [B,G,L] or NULL */
  float* y_soft;             /* [B,G,L] */
  float* soft;               /* [B,G,L] or NULL */
  int32_t* idx;              /* [B,L] */
  float* count;              /* [B,G] */
} sc_assign_desc;
int sc_assign_fwd(const sc_assign_desc* a, void* stream);

/* agg[b,g,:] = sum_{l: idx=g} v[b,l,:] / max(count[b,g], 1) (:309-310); sum_out = qf + agg (:312). */
int sc_aggregate_fwd(const void* v, int v_dtype, const int32_t* idx, const float* count, const float* qf, float* agg,
                     float* sum_out, int B, int L, int D, void* stream);

/* sc_assign_fwd + sc_aggregate_fwd as ONE kernel (one CTA per sample; the sample's key and value tiles are each read once,
 * the patch's centre and the counts never leave shared memory).  `count` is written, not accumulated.  Same outputs as the
 * two calls; shapes the fused kernel does not take fall back to them inside the call. */
int sc_assign_aggregate_fwd(const sc_assign_desc* a, const void* v, int v_dtype, float* agg, float* sum_out, void* stream);

/* Backward of the three steps above given d_agg = d(sum_out):
 *   d hard[g,l] = (<d_agg_g, v_l> - [count_g >= 1] <d_agg_g, agg_g>) / max(count_g,1) + d_hard_extra[g,l]
 *   d logits    = y_soft * (d hard - sum_c y_soft_c d hard_c) / tau           (straight-through, :237)
 *   d v[l,:]    = d_agg[idx_l,:] / max(count_idx_l, 1)
 *   d k[l,:]    = sum_g d logits[g,l] qf[g,:];   d_qf = d_qf_base + sum_l d logits[g,l] k[l,:]
 * d_hard_extra carries the superpixel-KL and ReconstructLayer gradients w.r.t. hard_attn. */
typedef struct {
  int32_t B, G, L, D;
  const float* d_agg; /* [B,G,D] */
  const float* agg;
  const void* v;
  int32_t v_dtype;
  const int32_t* idx;
  const float* count;
  const float* y_soft;
  const float* d_hard_extra; /* [B,G,L] or NULL */
  float tau;
  const float* qf;
  const void* k;
  int32_t k_dtype;
  float* d_logits;         /* [B,G,L] scratch/out */
  void* d_v;               /* [B,L,D] v_dtype */
  void* d_k;               /* [B,L,D] k_dtype */
  const float* d_qf_base;  /* [B,G,D] or NULL */
  float* d_qf;             /* [B,G,D] */
} sc_assign_bwd_desc;
int sc_assign_bwd(const sc_assign_bwd_desc* a, void* stream);

/* ReconstructLayer (modules/module_seg_vit.py:333-345) with a one-hot assignment:
 *   pre[b,m,:] = sum_g' (W[g', idx[b,m]] + bias[g']) sx[b,g',:];  out = QuickGELU(pre)
 * backward: d_sx, d hard[b,g,m] (into the d_hard_extra of sc_assign_bwd), dW / dbias (atomic +=). */
int sc_reconstruct_fwd(const float* sx, const int32_t* idx, const float* W, const float* bias, float* pre, float* out,
                       int B, int M, int D, void* stream);
int sc_reconstruct_bwd(const float* d_out, const float* pre, const float* sx, const int32_t* idx, const float* W,
                       const float* bias, float* d_sx, float* d_hard, float* dW, float* dbias, int B, int M, int D,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Loss heads (modules/modeling.py:201-252).  Every loss term is atomically added into one device
 * scalar `loss`; `gscale` is the upstream gradient (d total / d loss, normally 1).
 */
/* y = x / ||x||_2 per row (modeling.py:341-345) and backward. */
int sc_l2norm_fwd(const float* x, float* y, float* inv_norm, int rows, int E, void* stream);
int sc_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int rows, int E, void* stream);

/* InfoNCE, one direction (modeling.py:204-209,348-357).  raw = X_loc Y_all^T [B, N] (cosines, N = B*W);
 * s = min(exp(*logit_scale_param), 100); labels = label_off + i with label_off = B*rank.
 *   sc_ce_lse : lse[i] = logsumexp_j(s raw_ij);  loss += (lse_i - s raw_i,label) * 0.5 / B
 *   sc_ce_grad: raw_ij <- s * gscale * 0.5/B * [ (softmax_own_ij - 1{j=label}) +
 *                                                 (exp(s raw_ij - lse_other_all[j]) - 1{j=label}) ]
 *     i.e. d loss / d (X_loc . Y_all) including the column terms that the reference obtains through
 *     diffdist's reduce-scatter (modules/util_module.py:180-190) -- lse_other_all is the all-gathered
 *     lse of the opposite direction, so no reduction collective is needed;
 *     *d_logit_scale_param += sum_ij own-row gradient * raw_ij * d s/d p   (own rows only, like the reference). */
int sc_ce_lse(const float* raw, int B, int N, int label_off, const float* logit_scale_param, float* lse, float* loss,
              void* stream);
int sc_ce_grad(float* raw, int B, int N, int label_off, const float* logit_scale_param, const float* lse_own,
               const float* lse_other_all, float gscale, float* d_logit_scale_param, void* stream);

/* Superpixel-KL head (modeling.py:212-224) on the hard assignment idx [B,L] and labels seg [B,L] (int64):
 * loss += symmetric KL / (2 B L G);  d_hard[b,g,l] = d loss / d hard_attn[b,g,l] * gscale. */
int sc_superpixel_kl(const int32_t* idx, const int64_t* seg, int B, int L, float gscale, float* loss, float* d_hard,
                     void* stream);

/* MAE decoder input assembly (modules/module_mae.py:306-311): x[b,i,:] = (r = ids_restore[b,i]) < keep ?
 * emb[b,r,:] : mask_token, plus decoder_pos_embed[i,:]; and its backward. */
int sc_mae_unshuffle(const void* emb, int emb_dtype, const float* mask_token, const int32_t* ids_restore, const float* pos,
                     float* x, int B, int L1, int keep, int D, void* stream);
int sc_mae_unshuffle_bwd(const float* dx, const int32_t* ids_restore, void* d_emb, int emb_dtype, float* d_mask_token, int B,
                         int L1, int keep, int D, void* stream);
/* Masked-patch MSE against patchify(image) (module_mae.py:18-29,322-328): loss += ...; dpred written for all
 * L1 rows (zero for CLS / kept patches). */
int sc_mae_loss(const void* pred, int dtype, const float* image, const float* mask, int B, int L1, int keep, int grid,
                int patch, float gscale, float* loss, void* dpred, void* stream);

/* ------------------------------------------------------------------------------------------------
 * NVLink peer-to-peer all-gather of the contrastive head (replaces diffdist.functional.all_gather +
 * torch.distributed.barrier(), modules/util_module.py:180-190 and modules/modeling.py:352-354).
 * Set-up (once): every rank sc_p2p_alloc()s a data buffer + signal pad, the 64-byte IPC handles are exchanged
 * out of band (torch.distributed all_gather_object) and opened with sc_p2p_open().
 * Hot path: sc_p2p_allgather copies this rank's segments (host arrays srcs/nbytes/offs, <= 4) into the same byte
 * offsets of EVERY peer's buffer with plain stores
 * over NVLink, raises per-peer epoch flags (st.release.sys) and waits for the peers' flags (ld.acquire.sys);
 * sc_p2p_release tells the peers that this rank is done reading epoch `epoch` (so they may overwrite).
 * epoch must increase by 1 per exchange on every rank.  scratch: device uint32[2], zeroed once.
 */
int sc_p2p_alloc(int64_t bytes, void** buf, void** pad, void* buf_handle_out, void* pad_handle_out);
int sc_p2p_open(const void* handle, void** ptr);
int sc_p2p_close(void* ptr);
int sc_p2p_free(void* buf, void* pad);
int sc_p2p_allgather(int nseg, const void* const* srcs, const int64_t* nbytes, const int64_t* offs, void* const* peer_bufs,
                     void* const* peer_pads, int rank, int world, uint32_t epoch, void* scratch, void* stream);
int sc_p2p_release(void* const* peer_pads, int rank, int world, uint32_t epoch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" row: fused optimizer step = torch.nn.utils.clip_grad_norm_(params, max_norm)
 * (main_task_align.py:326) + AdaptAdamW.step (modules/optimization_adamw.py:112-174: decoupled weight decay BEFORE the
 * Adam update, bias-corrected step size and denominator) + the logit_scale clamp (main_task_align.py:343-347).
 * One table entry per parameter tensor; per-tensor scalars are computed on the host from the group's step count and
 * cosine-warm-up schedule:  decay = 1 - lr*wd,  step_size = lr / (1 - b1^t),  inv_sqrt_bc2 = 1 / sqrt(1 - b2^t).
 *   sc_grad_sqnorm_multi: *out_sqnorm += sum over all tensors of ||grad||^2     (caller zeroes out_sqnorm)
 *   sc_adamw_multi      : g' = g * min(1, max_norm/(sqrt(*sqnorm)+1e-6));  m,v update;  p = p*decay - step_size*m/(sqrt(v)*inv_sqrt_bc2+eps)
 */
typedef struct {
  void* param;        /* fp32, updated in place */
  const void* grad;   /* fp32 or NULL (tensor skipped, like `if p.grad is None: continue`) */
  void* exp_avg;      /* fp32 state */
  void* exp_avg_sq;   /* fp32 state */
  int64_t n;
  int64_t first_block; /* prefix sum of ceil(n / 1024) */
  float decay;
  float step_size;
  float inv_sqrt_bc2;
  float clamp_max;
  int32_t clamp_max_enabled;
  int32_t pad_;
} sc_opt_item;
int sc_grad_sqnorm_multi(const sc_opt_item* items_dev, int n_items, int64_t total_blocks, float* out_sqnorm, void* stream);
int sc_adamw_multi(const sc_opt_item* items_dev, int n_items, int64_t total_blocks, const float* sqnorm, float max_norm,
                   float beta1, float beta2, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient mean over NVSwitch multicast (replaces the DDP bucket all-reduce,
 * main_task_align.py:251-252).  `multicast_ptr` is the NVLS multicast address of `count` fp32 values that live at the
 * same offset of a symmetric buffer on every rank; after the call every replica holds scale * sum over ranks.
 * peer_flags_dev: device array of `world` pointers, entry t = rank t's flag array (uint32 [SC_NVLS_MAX_BLOCKS * world],
 * zero-initialised, peer-mapped).  `epoch` must advance by 2 per call, identically on every rank.  *err_flag is set to 1
 * if a peer does not arrive within timeout_ms (the kernel then returns instead of hanging).
 */
#define SC_NVLS_MAX_BLOCKS 64
int sc_nvls_allreduce(void* multicast_ptr, int64_t count, float scale, int rank, int world, void* const* peer_flags_dev,
                      uint32_t epoch, int blocks, int64_t timeout_ms, int32_t* err_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGCLIP_B200_H */
