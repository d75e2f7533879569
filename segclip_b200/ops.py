"""Thin tensor-level wrappers over the C ABI (PyTorch tensors in, PyTorch tensors out).
Only pointer/shape marshalling happens here -- no arithmetic."""
import ctypes as C

import torch

from . import _lib as L
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_QUICKGELU, BF16, F32  # noqa: F401


def _chk2d(t):
    assert t.dim() == 2 and t.stride(1) == 1 and t.is_cuda, (t.shape, t.stride())


def gemm_desc(A, B, C_, *, trans_a=False, trans_b=False, bias=None, rowbias=None, rowbias_idx=None, rowbias_mod=0,
              act=ACT_NONE, residual=None, C2=None, accumulate=False, split_k=0, alpha=1.0, force_simt=False):
    """Builds the sc_gemm descriptor for C = epi(A B^T); see include/segclip_b200.h."""
    for t in (A, B, C_):
        _chk2d(t)
    M, N = C_.shape
    K = A.shape[0] if trans_a else A.shape[1]
    assert (A.shape[1] if trans_a else A.shape[0]) == M, (A.shape, C_.shape, trans_a)
    assert (B.shape == (K, N)) if trans_b else (B.shape == (N, K)), (B.shape, N, K, trans_b)
    assert A.dtype == B.dtype
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
        assert residual.dtype == torch.float32 and residual.shape == C_.shape
        d.residual, d.ldr = residual.data_ptr(), residual.stride(0)
    d.C, d.ldc, d.c_dtype = C_.data_ptr(), C_.stride(0), L.dt(C_)
    if C2 is not None:
        _chk2d(C2)
        assert C2.shape == C_.shape and C2.stride(0) == C_.stride(0)
        d.C2, d.c2_dtype = C2.data_ptr(), L.dt(C2)
    d.accumulate, d.split_k, d.force_simt = int(accumulate), split_k, int(force_simt)
    return d


def gemm(A, B, C_, **kw):
    d = gemm_desc(A, B, C_, **kw)
    L.check(L.lib().sc_gemm(C.byref(d), L.stream()), "sc_gemm")
    return C_
