import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import ops
M, D = 50176, 768
x = torch.randn(M, D, device="cuda"); dy = torch.randn(M, D, device="cuda").bfloat16()
g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
dx = torch.randn(M, D, device="cuda"); dxT = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
f = ops.layernorm_op(x, g, b, y, 1e-5, mean, rstd)
bw = ops.layernorm_bwd_op(dy, x, mean, rstd, g, dx, True, dxT, dg, db)
cs = ops.colsum_op(torch.randn(M, 3072, device="cuda").bfloat16(), torch.zeros(3072, device="cuda"))
for name, op, nbytes in (("ln_fwd", f, M * D * 6), ("ln_bwd", bw, M * D * 16), ("colsum3072", cs, M * 3072 * 2)):
    for _ in range(3): op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): op()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-10s %.3f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6))
bw2 = ops.layernorm_bwd_op(dy, x, mean, rstd, g, dx, True, dxT, None, None)
bw3 = ops.layernorm_bwd_op(dy, x, mean, rstd, g, dx, False, None, dg, db)
for name, op, nbytes in (("ln_bwd_noparam", bw2, M * D * 16), ("ln_bwd_noacc_nocopy", bw3, M * D * 10)):
    for _ in range(3): op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): op()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-20s %.3f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6))
