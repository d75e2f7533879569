"""Micro-benchmark of sc_gemm (tcgen05) on the hot-path problem sizes; prints TFLOP/s per shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import ops  # noqa: E402

SHAPES = [  # name, M, N, K, ta, tb, kwargs
    ("qkv      fwd", 50176, 2304, 768, False, False, dict(bias=True, out=torch.bfloat16)),
    ("out_proj fwd", 50176, 768, 768, False, False, dict(bias=True, residual=True, out=torch.float32)),
    ("c_fc     fwd", 50176, 3072, 768, False, False, dict(bias=True, act=1, c2=True, out=torch.bfloat16)),
    ("c_proj   fwd", 50176, 768, 3072, False, False, dict(bias=True, residual=True, out=torch.float32)),
    ("c_proj dgrad", 50176, 3072, 768, False, True, dict(out=torch.bfloat16)),
    ("c_proj dgrad*", 50176, 3072, 768, False, True, dict(out=torch.bfloat16, aux=True)),
    ("c_fc   dgrad", 50176, 768, 3072, False, True, dict(out=torch.bfloat16)),
    ("c_fc   wgrad", 3072, 768, 50176, True, True, dict(acc=True, out=torch.float32)),
    ("c_proj wgrad", 768, 3072, 50176, True, True, dict(acc=True, out=torch.float32)),
    ("qkv    wgrad", 2304, 768, 50176, True, True, dict(acc=True, out=torch.float32)),
    ("text qkv fwd", 19712, 1536, 512, False, False, dict(bias=True, out=torch.bfloat16)),
    ("text c_fc fwd", 19712, 2048, 512, False, False, dict(bias=True, act=1, c2=True, out=torch.bfloat16)),
    ("text c_proj dgrad*", 19712, 2048, 512, False, True, dict(out=torch.bfloat16, aux=True)),
    ("text c_proj fwd", 19712, 512, 2048, False, False, dict(bias=True, residual=True, out=torch.float32)),
    ("text out_proj fwd", 19712, 512, 512, False, False, dict(bias=True, residual=True, out=torch.float32)),
    ("c_proj dgrad*+cs", 50176, 3072, 768, False, True, dict(out=torch.bfloat16, aux=True, colsum=True)),
    ("c_fc fwd (deriv)", 50176, 3072, 768, False, False, dict(bias=True, act=1, c2=True, out=torch.bfloat16, deriv=True)),
    ("c_proj dgrad' +cs", 50176, 3072, 768, False, True, dict(out=torch.bfloat16, aux=True, colsum=True, deriv=True)),
    ("text c_proj dgrad'", 19712, 2048, 512, False, True, dict(out=torch.bfloat16, aux=True, colsum=True, deriv=True)),
    ("plain 8192^3", 8192, 8192, 8192, False, False, dict(out=torch.bfloat16)),
]


def main():
    dev = "cuda"
    only = sys.argv[1] if len(sys.argv) > 1 else None       # substring filter (ncu captures of one shape)
    for name, M, N, K, ta, tb, kw in SHAPES:
        if only and only not in name:
            continue
        A = torch.randn((K, M) if ta else (M, K), device=dev).bfloat16()
        B = torch.randn((K, N) if tb else (N, K), device=dev).bfloat16()
        C = torch.zeros(M, N, device=dev, dtype=kw["out"])
        bias = torch.randn(N, device=dev) if kw.get("bias") else None
        res = torch.randn(M, N, device=dev) if kw.get("residual") else None
        C2 = torch.empty(M, N, device=dev, dtype=kw["out"]) if kw.get("c2") else None
        aux = torch.randn(M, N, device=dev).bfloat16() if kw.get("aux") else None
        cs = torch.zeros(N, device=dev) if kw.get("colsum") else None
        op = ops.gemm_op(A, B, C, trans_a=ta, trans_b=tb, bias=bias, residual=res, act=kw.get("act", 0), C2=C2,
                         mul_aux=aux, mul_aux_act=(3 if kw.get("deriv") else 1) if aux is not None else 0, colsum_out=cs,
                         c2_is_act_grad=bool(kw.get("deriv")) and C2 is not None,
                         accumulate=kw.get("acc", False), split_k=-1 if kw.get("acc") else 0)
        for _ in range(3):
            op()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            op()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("%-18s M=%6d N=%5d K=%6d  %8.3f ms  %7.1f TFLOP/s" % (name, M, N, K, ms, 2.0 * M * N * K / ms / 1e9))


if __name__ == "__main__":
    main()
