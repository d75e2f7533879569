"""TEST INFRASTRUCTURE ONLY -- "ideal bf16" yardstick for the gradient tolerance of the CUDA path.

The CUDA path computes every matrix product with bf16 operands and fp32 accumulation (tcgen05 kind::f16), forward and
backward.  Even a perfect implementation of that arithmetic differs from the fp32 oracle: operand rounding alone moves
the gradients by several percent through 12+ layers.  ``BF16Operands`` re-runs the UNCHANGED fp32 oracle
(oracle/segclip_oracle.py) under a TorchFunctionMode that rounds the tensor operands of every matrix product
(F.linear, matmul / @, einsum, conv2d) to bf16 and back, in the forward pass and -- through a gradient hook on each
product's output -- the incoming gradient of the two backward products as well.  Everything else (LayerNorm, softmax,
GELU, residual adds, accumulation) stays fp32, exactly like the CUDA path's fp32 epilogues.

    with BF16Operands():
        loss, grads, info = so.loss_and_grads(...)

The difference between those gradients and the plain fp32 oracle's is the irreducible share of the bf16 error; the GPU
parity tests bound the CUDA path's error by a small multiple of it (tests/test_e2e_gpu.py).
"""
import torch
import torch.nn.functional as F
from torch.overrides import TorchFunctionMode


class _RoundFwd(torch.autograd.Function):
    """x -> bf16(x) as fp32; gradient passes unchanged."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """identity; the incoming gradient is rounded to bf16 (it is the operand of the dgrad / wgrad products)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def _r(t):
    if isinstance(t, torch.Tensor) and t.is_floating_point() and t.dtype == torch.float32:
        return _RoundFwd.apply(t)
    return t


class BF16Operands(TorchFunctionMode):
    PRODUCTS = {F.linear, torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__, torch.einsum, F.conv2d, torch.mm,
                torch.bmm, torch.Tensor.mm, torch.Tensor.bmm}

    def __init__(self):
        super().__init__()
        self.products = 0

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func not in self.PRODUCTS:
            return func(*args, **kwargs)
        self.products += 1
        if func is F.linear:                      # the bias is added in fp32 by the epilogue: not rounded
            x, w = _r(args[0]), _r(args[1])
            rest = args[2:]
            out = func(x, w, *rest, **kwargs)
        elif func is F.conv2d:
            out = func(_r(args[0]), _r(args[1]), *args[2:], **kwargs)
        elif func is torch.einsum:
            out = func(args[0], *[_r(a) for a in args[1:]], **kwargs)
        else:
            out = func(*[_r(a) for a in args], **kwargs)
        return _RoundBwd.apply(out) if out.requires_grad else out
