"""Micro-benchmark of the LayerNorm kernels and the column-sum kernel on the hot-path sizes (GB/s of algorithmic bytes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import ops


def timeit(name, op, nbytes):
    for _ in range(3): op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): op()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-34s %.3f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6))


for M, D in ((50176, 768), (19712, 512)):
    x = torch.randn(M, D, device="cuda"); dy = torch.randn(M, D, device="cuda").bfloat16()
    g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
    y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    dx = torch.randn(M, D, device="cuda"); dxT = torch.randn(M, D, device="cuda").bfloat16()
    dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda"); cs = torch.zeros(D, device="cuda")
    print("# rows=%d D=%d" % (M, D))
    timeit("ln_fwd f32->bf16", ops.layernorm_op(x, g, b, y, 1e-5, mean, rstd), M * D * 6)
    timeit("ln_bwd f32 stream + bf16 twin", ops.layernorm_bwd_op(dy, x, mean, rstd, g, dx, True, dxT, dg, db), M * D * 16)
    timeit("ln_bwd bf16 stream", ops.layernorm_bwd_op(dy, x, mean, rstd, g, dxT, True, None, dg, db), M * D * 10)
    timeit("ln_bwd bf16 stream + colsum", ops.layernorm_bwd_op(dy, x, mean, rstd, g, dxT, True, None, dg, db, dx_colsum=cs), M * D * 10)
    timeit("ln_bwd bf16 stream, no params", ops.layernorm_bwd_op(dy, x, mean, rstd, g, dxT, True, None, None, None), M * D * 10)
    timeit("colsum D (bf16)", ops.colsum_op(dxT, cs), M * D * 2)
    big = torch.randn(M, 3 * D, device="cuda").bfloat16(); cs3 = torch.zeros(3 * D, device="cuda")
    timeit("colsum 3D (bf16)", ops.colsum_op(big, cs3), M * 3 * D * 2)
