"""Runs one hot-path GEMM a few times (for ncu captures).  usage: one_gemm.py [c_fc|qkv|out_proj|plain]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c_fc"
M = 50176
dev = "cuda"
if which == "c_fc":
    N, K = 3072, 768
    kw = dict(bias=torch.randn(N, device=dev), act=1, C2=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
elif which == "qkv":
    N, K = 2304, 768
    kw = dict(bias=torch.randn(N, device=dev))
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
elif which == "out_proj":
    N, K = 768, 768
    kw = dict(bias=torch.randn(N, device=dev), residual=torch.randn(M, N, device=dev))
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
else:
    M, N, K = 8192, 8192, 8192
    kw = {}
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
A = torch.randn(M, K, device=dev).bfloat16()
B = torch.randn(N, K, device=dev).bfloat16()
op = ops.gemm_op(A, B, out, **kw)
for _ in range(5):
    op()
torch.cuda.synchronize()
