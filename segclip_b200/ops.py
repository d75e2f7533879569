"""Descriptor builders over the C ABI (PyTorch tensors in, prepared native calls out).

Every function returns an ``Op`` -- (C function, argument tuple, tensors kept alive) -- that the
engine replays each step with only the CUDA stream appended.  No arithmetic happens in Python."""
import ctypes as C
import os

import torch

from . import _lib as L
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_QUICKGELU, BF16, F32  # noqa: F401

ACT_DERIV = 3        # mul_aux_act: mul_aux already holds act'(pre-activation) (forward gemm with c2_is_act_grad)


class Op:
    __slots__ = ("fn", "args", "keep", "name")

    def __init__(self, name, args, keep=()):
        self.name = name
        self.fn = getattr(L.lib(), name)
        self.args = tuple(args)
        self.keep = keep

    def __call__(self, stream=None):
        rc = self.fn(*self.args, L.stream() if stream is None else stream)
        if rc != 0:
            L.check(rc, self.name)


def _chk2d(t):
    assert t.dim() == 2 and t.stride(1) == 1 and t.is_cuda, (t.shape, t.stride())


def _p(t):
    return None if t is None else t.data_ptr()


def gemm_desc(A, B, C_, *, trans_a=False, trans_b=False, bias=None, rowbias=None, rowbias_idx=None, rowbias_mod=0,
              act=ACT_NONE, residual=None, C2=None, accumulate=False, split_k=0, alpha=1.0, force_simt=False,
              mul_aux=None, mul_aux_act=ACT_NONE, colsum_out=None, dot_aux=None, dot_out=None, dot_L=0, c2_is_act_grad=False):
    """sc_gemm descriptor for C = epi(A B^T); see include/segclip_b200.h."""
    for t in (A, B, C_):
        _chk2d(t)
    M, N = C_.shape
    K = A.shape[0] if trans_a else A.shape[1]
    assert (A.shape[1] if trans_a else A.shape[0]) == M, (A.shape, C_.shape, trans_a)
    assert (tuple(B.shape) == (K, N)) if trans_b else (tuple(B.shape) == (N, K)), (B.shape, N, K, trans_b)
    assert A.dtype == B.dtype, (A.dtype, B.dtype)
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.in_dtype = L.dt(A)
    d.trans_a, d.trans_b = int(trans_a), int(trans_b)
    d.A, d.lda, d.B, d.ldb = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
    d.alpha = alpha
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
        d.bias = bias.data_ptr()
    if rowbias is not None:
        _chk2d(rowbias)
        assert rowbias.dtype == torch.float32 and rowbias.shape[1] == N
        d.rowbias, d.ld_rowbias = rowbias.data_ptr(), rowbias.stride(0)
        if rowbias_idx is not None:
            assert rowbias_idx.dtype == torch.int32 and rowbias_idx.numel() == M
            d.rowbias_idx = rowbias_idx.data_ptr()
        d.rowbias_mod = rowbias_mod if rowbias_mod else rowbias.shape[0]
    d.act = act
    if residual is not None:
        _chk2d(residual)
        assert residual.shape == C_.shape and (residual.dtype == torch.float32 or residual.dtype == C_.dtype), (residual.dtype, C_.dtype)
        d.residual, d.ldr, d.residual_dtype = residual.data_ptr(), residual.stride(0), L.dt(residual)
    d.C, d.ldc, d.c_dtype = C_.data_ptr(), C_.stride(0), L.dt(C_)
    if C2 is not None:
        _chk2d(C2)
        assert C2.shape == C_.shape and C2.stride(0) == C_.stride(0)
        d.C2, d.c2_dtype = C2.data_ptr(), L.dt(C2)
        d.c2_is_act_grad = int(c2_is_act_grad)
    if accumulate:
        assert C_.dtype == torch.float32
    d.accumulate, d.split_k, d.force_simt = int(accumulate), split_k, int(force_simt)
    if mul_aux is not None:
        _chk2d(mul_aux)
        assert mul_aux.shape == C_.shape and mul_aux.stride(0) == C_.stride(0)
        d.mul_aux, d.mul_aux_dtype, d.mul_aux_act = mul_aux.data_ptr(), L.dt(mul_aux), mul_aux_act
    if colsum_out is not None:
        assert colsum_out.dtype == torch.float32 and colsum_out.numel() == N and C_.dtype == torch.bfloat16
        d.colsum_out = colsum_out.data_ptr()
    if dot_out is not None:
        _chk2d(dot_aux)
        assert dot_aux.shape == C_.shape and dot_aux.stride(0) == C_.stride(0) and dot_aux.dtype == C_.dtype == torch.bfloat16
        assert dot_out.dtype == torch.float32 and dot_out.numel() == M * (N // 64) and dot_L > 0 and M % dot_L == 0
        d.dot_aux, d.dot_out, d.dot_L = dot_aux.data_ptr(), dot_out.data_ptr(), dot_L
    return d


def gemm_op(A, B, C_, **kw):
    d = gemm_desc(A, B, C_, **kw)
    return Op("sc_gemm", (C.byref(d),), (d, A, B, C_, kw))


def gemm(A, B, C_, **kw):
    gemm_op(A, B, C_, **kw)()
    return C_


def layernorm_op(x, gamma, beta, y, eps=1e-5, mean=None, rstd=None, remap=None):
    rows, D = x.shape
    d = L.LnDesc()
    d.rows, d.D = rows, D
    d.x, d.x_dtype = x.data_ptr(), L.dt(x)
    d.gamma, d.beta, d.eps = gamma.data_ptr(), beta.data_ptr(), eps
    d.y, d.y_dtype = y.data_ptr(), L.dt(y)
    if remap is not None:
        d.in_group, d.out_group, d.out_off = remap
    d.mean, d.rstd = _p(mean), _p(rstd)
    assert x.is_contiguous() and y.is_contiguous()
    return Op("sc_layernorm_fwd", (C.byref(d),), (d, x, gamma, beta, y, mean, rstd))


def layernorm_bwd_op(dy, x, mean, rstd, gamma, dx=None, accumulate_dx=False, dx_copy=None, dgamma=None, dbeta=None,
                     remap=None, dx_colsum=None):
    rows, D = x.shape
    d = L.LnBwdDesc()
    d.rows, d.D = rows, D
    d.dy, d.dy_dtype = dy.data_ptr(), L.dt(dy)
    d.x, d.x_dtype = x.data_ptr(), L.dt(x)
    d.mean, d.rstd, d.gamma = mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr()
    if dx is not None:
        d.dx, d.dx_dtype = dx.data_ptr(), L.dt(dx)
    d.accumulate_dx = int(accumulate_dx)
    if dx_copy is not None:
        assert dx_copy.dtype == torch.bfloat16 and dx is not None
        d.dx_copy_bf16 = dx_copy.data_ptr()
    d.dgamma, d.dbeta = _p(dgamma), _p(dbeta)
    if remap is not None:
        d.in_group, d.out_group, d.out_off = remap
    if dx_colsum is not None:
        assert dx_colsum.dtype == torch.float32 and dx_colsum.numel() == D
        d.dx_colsum = dx_colsum.data_ptr()
    return Op("sc_layernorm_bwd", (C.byref(d),), (d, dy, x, mean, rstd, gamma, dx, dx_copy, dgamma, dbeta, dx_colsum))


def attn_desc(q, k, v, o, lse, B, H, Lq, Lk, hd, q_str, k_str, v_str, o_str, causal=False, force_generic=False):
    """q/k/v/o are base tensors (already offset to their first element); *_str = (batch stride, row stride)."""
    a = L.AttnDesc()
    a.B, a.H, a.Lq, a.Lk, a.hd = B, H, Lq, Lk, hd
    a.dtype = L.dt(q)
    a.causal = int(causal)
    a.scale = hd ** -0.5
    a.q, (a.q_bs, a.q_rs) = q.data_ptr(), q_str
    a.k, (a.k_bs, a.k_rs) = k.data_ptr(), k_str
    a.v, (a.v_bs, a.v_rs) = v.data_ptr(), v_str
    a.o, (a.o_bs, a.o_rs) = o.data_ptr(), o_str
    a.lse = lse.data_ptr()
    a.force_generic = int(force_generic or bool(os.environ.get("SC_ATT_GENERIC")))
    return a


def attention_op(a, keep=()):
    return Op("sc_attention_fwd", (C.byref(a),), (a, keep))


def attention_bwd_op(a, d_o, d_q, d_k, d_v, delta_ws=None, keep=(), bias_grad=None, delta_ready=False):
    """bias_grad: optional fp32 [3*H*hd] gradient of the fused q|k|v projection bias (+= column sums of d_q, d_k, d_v);
    delta_ready: delta_ws already holds rowsum(dO o O) (gemm_op(..., dot_out=delta_ws))."""
    g = L.AttnBwdDesc()
    g.fwd = a
    g.d_o, g.d_q, g.d_k, g.d_v = d_o.data_ptr(), d_q.data_ptr(), d_k.data_ptr(), d_v.data_ptr()
    if bias_grad is not None:
        W = a.H * a.hd
        assert bias_grad.dtype == torch.float32 and bias_grad.numel() == 3 * W and bias_grad.is_contiguous()
        g.dq_colsum, g.dk_colsum, g.dv_colsum = bias_grad.data_ptr(), bias_grad.data_ptr() + 4 * W, bias_grad.data_ptr() + 8 * W
        keep = (keep, bias_grad)
    if delta_ws is not None:
        assert delta_ws.dtype == torch.float32 and delta_ws.numel() >= a.B * a.H * a.Lq
        g.delta_ws = delta_ws.data_ptr()
        g.delta_ready = int(delta_ready)
    return Op("sc_attention_bwd", (C.byref(g),), (g, d_o, d_q, d_k, d_v, delta_ws, keep))


def act_bwd_op(dy, pre, dx, act):
    assert dy.numel() == pre.numel() == dx.numel()
    return Op("sc_act_bwd", (dy.data_ptr(), L.dt(dy), pre.data_ptr(), L.dt(pre), dx.data_ptr(), L.dt(dx), dy.numel(), act),
              (dy, pre, dx))


def convert_op(src, dst, scale=None):
    assert src.numel() == dst.numel()
    return Op("sc_convert", (src.data_ptr(), L.dt(src), dst.data_ptr(), L.dt(dst), src.numel(), _p(scale)), (src, dst, scale))


def colsum_op(x, out, rows=None, cols=None, ld=None):
    """out[c] += sum_r x[r, c]; x may be viewed as [rows, cols] with leading dimension ld."""
    if rows is None:
        _chk2d(x)
        rows, cols, ld = x.shape[0], x.shape[1], x.stride(0)
    assert out.dtype == torch.float32 and out.numel() == cols
    return Op("sc_colsum", (x.data_ptr(), L.dt(x), ld, rows, cols, out.data_ptr()), (x, out))


def cast_multi_op(items_dev, n_items, total_blocks, dst_dtype, keep=()):
    return Op("sc_cast_multi", (items_dev.data_ptr(), n_items, total_blocks, dst_dtype), (items_dev, keep))


def blockdiag_expand_op(w, dense, groups):
    Cc = dense.shape[0]
    return Op("sc_blockdiag_expand", (w.data_ptr(), dense.data_ptr(), Cc, groups, L.dt(dense)), (w, dense))


def blockdiag_reduce_op(dense_grad, dw, groups):
    Cc = dense_grad.shape[0]
    return Op("sc_blockdiag_reduce", (dense_grad.data_ptr(), dw.data_ptr(), Cc, groups), (dense_grad, dw))


def im2col_op(image, out, patch_idx, rows_per_img, grid_hw, patch):
    gh, gw = grid_hw
    return Op("sc_im2col", (image.data_ptr(), out.data_ptr(), L.dt(out), out.stride(0), _p(patch_idx), out.shape[0],
                            rows_per_img, gh, gw, patch), (image, out, patch_idx))


def bicubic_resize_op(src, dst, src_hw, dst_hw):
    """[sh*sw, D] -> [dh*dw, D] fp32 (positional table, inference at another input size)."""
    assert src.dtype == dst.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    return Op("sc_bicubic_resize", (src.data_ptr(), dst.data_ptr(), src_hw[0], src_hw[1], dst_hw[0], dst_hw[1], src.shape[-1]),
              (src, dst))


def text_embed_op(ids, tok, pos, out, eot_rows, B, T, W):
    return Op("sc_text_embed", (ids.data_ptr(), tok.data_ptr(), pos.data_ptr(), out.data_ptr(), eot_rows.data_ptr(), B, T, W),
              (ids, tok, pos, out, eot_rows))


def gather_rows_op(src, idx, out):
    assert out.dtype == torch.float32
    return Op("sc_gather_rows", (src.data_ptr(), L.dt(src), idx.data_ptr(), out.data_ptr(), out.shape[0], out.shape[1]), (src, idx, out))


def scatter_rows_op(src, idx, out):
    return Op("sc_scatter_rows", (src.data_ptr(), idx.data_ptr(), out.data_ptr(), src.shape[0], src.shape[1]), (src, idx, out))


def scatter_add_rows_op(src, idx, out, idx_offset=0):
    """out[idx[r] + idx_offset] += src[r] (embedding-lookup backward)."""
    assert idx.dtype in (torch.int32, torch.int64) and idx.numel() == src.shape[0] and out.dtype == torch.float32
    assert src.is_contiguous() and out.is_contiguous() and out.shape[-1] == src.shape[1]
    return Op("sc_scatter_add_rows", (src.data_ptr(), L.dt(src), idx.data_ptr(), int(idx.dtype == torch.int64), idx_offset,
                                      out.data_ptr(), src.shape[0], src.shape[1]), (src, idx, out))


def mae_mask_op(u, ids_restore, ids_keep, mask, patch_idx, B, L1, keep):
    return Op("sc_mae_mask", (u.data_ptr(), B, L1, keep, ids_restore.data_ptr(), ids_keep.data_ptr(), mask.data_ptr(),
                              patch_idx.data_ptr()), (u, ids_restore, ids_keep, mask, patch_idx))


def pool_max_op(x, out, arg, B, G, D):
    return Op("sc_pool_max", (x.data_ptr(), out.data_ptr(), arg.data_ptr(), B, G, D), (x, out, arg))


def pool_max_bwd_op(dout, arg, dx, B, G, D):
    return Op("sc_pool_max_bwd", (dout.data_ptr(), arg.data_ptr(), dx.data_ptr(), B, G, D), (dout, arg, dx))


def mean_cat_op(x, out, B, n, D):
    return Op("sc_mean_cat", (x.data_ptr(), out.data_ptr(), L.dt(out), B, n, D), (x, out))


def mean_cat_bwd_op(dout, dx, B, n, D):
    return Op("sc_mean_cat_bwd", (dout.data_ptr(), L.dt(dout), dx.data_ptr(), B, n, D), (dout, dx))


def assign_fwd_op(qf, k, u, y_soft, idx, count, B, Lp, D, tau=0.9, forced_idx=None, soft=None, logits=None):
    a = L.AssignDesc()
    a.B, a.G, a.L, a.D = B, 8, Lp, D
    a.qf, a.k, a.k_dtype, a.u, a.tau = qf.data_ptr(), k.data_ptr(), L.dt(k), _p(u), tau
    a.forced_idx, a.logits, a.soft = _p(forced_idx), _p(logits), _p(soft)
    a.y_soft, a.idx, a.count = y_soft.data_ptr(), idx.data_ptr(), count.data_ptr()
    return Op("sc_assign_fwd", (C.byref(a),), (a, qf, k, u, y_soft, idx, count, forced_idx, soft, logits))


def assign_aggregate_fwd_op(qf, k, u, y_soft, idx, count, v, agg, sum_out, B, Lp, D, tau=0.9, forced_idx=None, soft=None,
                            logits=None):
    """assign_fwd_op + aggregate_fwd_op as one kernel (sc_assign_aggregate_fwd)."""
    a = L.AssignDesc()
    a.B, a.G, a.L, a.D = B, 8, Lp, D
    a.qf, a.k, a.k_dtype, a.u, a.tau = qf.data_ptr(), k.data_ptr(), L.dt(k), _p(u), tau
    a.forced_idx, a.logits, a.soft = _p(forced_idx), _p(logits), _p(soft)
    a.y_soft, a.idx, a.count = y_soft.data_ptr(), idx.data_ptr(), count.data_ptr()
    return Op("sc_assign_aggregate_fwd", (C.byref(a), v.data_ptr(), L.dt(v), agg.data_ptr(), sum_out.data_ptr()),
              (a, qf, k, u, y_soft, idx, count, forced_idx, soft, logits, v, agg, sum_out))


def aggregate_fwd_op(v, idx, count, qf, agg, sum_out, B, Lp, D):
    return Op("sc_aggregate_fwd", (v.data_ptr(), L.dt(v), idx.data_ptr(), count.data_ptr(), qf.data_ptr(), agg.data_ptr(),
                                   sum_out.data_ptr(), B, Lp, D), (v, idx, count, qf, agg, sum_out))


def assign_bwd_op(d_agg, agg, v, idx, count, y_soft, d_hard_extra, qf, k, d_logits, d_v, d_k, d_qf_base, d_qf, B, Lp, D,
                  tau=0.9):
    a = L.AssignBwdDesc()
    a.B, a.G, a.L, a.D = B, 8, Lp, D
    a.d_agg, a.agg, a.v, a.v_dtype = d_agg.data_ptr(), agg.data_ptr(), v.data_ptr(), L.dt(v)
    a.idx, a.count, a.y_soft = idx.data_ptr(), count.data_ptr(), y_soft.data_ptr()
    a.d_hard_extra, a.tau = _p(d_hard_extra), tau
    a.qf, a.k, a.k_dtype = qf.data_ptr(), k.data_ptr(), L.dt(k)
    a.d_logits, a.d_v, a.d_k = d_logits.data_ptr(), d_v.data_ptr(), d_k.data_ptr()
    assert d_v.dtype == v.dtype and d_k.dtype == k.dtype
    a.d_qf_base, a.d_qf = _p(d_qf_base), d_qf.data_ptr()
    return Op("sc_assign_bwd", (C.byref(a),), (a, d_agg, agg, v, idx, count, y_soft, d_hard_extra, qf, k, d_logits, d_v, d_k,
                                               d_qf_base, d_qf))


def reconstruct_fwd_op(sx, idx, W, bias, pre, out, B, M, D):
    return Op("sc_reconstruct_fwd", (sx.data_ptr(), idx.data_ptr(), W.data_ptr(), bias.data_ptr(), pre.data_ptr(),
                                     out.data_ptr(), B, M, D), (sx, idx, W, bias, pre, out))


def reconstruct_bwd_op(d_out, pre, sx, idx, W, bias, d_sx, d_hard, dW, dbias, B, M, D):
    return Op("sc_reconstruct_bwd", (d_out.data_ptr(), pre.data_ptr(), sx.data_ptr(), idx.data_ptr(), W.data_ptr(),
                                     bias.data_ptr(), d_sx.data_ptr(), d_hard.data_ptr(), dW.data_ptr(), dbias.data_ptr(),
                                     B, M, D), (d_out, pre, sx, idx, W, bias, d_sx, d_hard, dW, dbias))


def l2norm_fwd_op(x, y, inv_norm):
    return Op("sc_l2norm_fwd", (x.data_ptr(), y.data_ptr(), inv_norm.data_ptr(), x.shape[0], x.shape[1]), (x, y, inv_norm))


def l2norm_bwd_op(dy, y, inv_norm, dx):
    return Op("sc_l2norm_bwd", (dy.data_ptr(), y.data_ptr(), inv_norm.data_ptr(), dx.data_ptr(), dy.shape[0], dy.shape[1]),
              (dy, y, inv_norm, dx))


def ce_lse_op(raw, label_off, scale_param, lse, loss):
    B, N = raw.shape
    return Op("sc_ce_lse", (raw.data_ptr(), B, N, label_off, scale_param.data_ptr(), lse.data_ptr(), loss.data_ptr()),
              (raw, scale_param, lse, loss))


def ce_grad_op(raw, label_off, scale_param, lse_own, lse_other_all, d_scale_param, gscale=1.0):
    B, N = raw.shape
    return Op("sc_ce_grad", (raw.data_ptr(), B, N, label_off, scale_param.data_ptr(), lse_own.data_ptr(),
                             lse_other_all.data_ptr(), gscale, d_scale_param.data_ptr()),
              (raw, scale_param, lse_own, lse_other_all, d_scale_param))


def superpixel_kl_op(idx, seg, loss, d_hard, B, Lp, gscale=1.0):
    return Op("sc_superpixel_kl", (idx.data_ptr(), seg.data_ptr(), B, Lp, gscale, loss.data_ptr(), d_hard.data_ptr()),
              (idx, seg, loss, d_hard))


def mae_unshuffle_op(emb, mask_token, ids_restore, pos, x, B, L1, keep, D):
    return Op("sc_mae_unshuffle", (emb.data_ptr(), L.dt(emb), mask_token.data_ptr(), ids_restore.data_ptr(), pos.data_ptr(),
                                   x.data_ptr(), B, L1, keep, D), (emb, mask_token, ids_restore, pos, x))


def mae_unshuffle_bwd_op(dx, ids_restore, d_emb, d_mask_token, B, L1, keep, D):
    return Op("sc_mae_unshuffle_bwd", (dx.data_ptr(), ids_restore.data_ptr(), d_emb.data_ptr(), L.dt(d_emb),
                                       d_mask_token.data_ptr(), B, L1, keep, D), (dx, ids_restore, d_emb, d_mask_token))


def mae_loss_op(pred, image, mask, loss, dpred, B, L1, keep, grid, patch, gscale=1.0):
    return Op("sc_mae_loss", (pred.data_ptr(), L.dt(pred), image.data_ptr(), mask.data_ptr(), B, L1, keep, grid, patch,
                              gscale, loss.data_ptr(), dpred.data_ptr()), (pred, image, mask, loss, dpred))
