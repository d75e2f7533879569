"""Debug: run one golden case twice (generic fp32-math attention vs tensor-core attention) and list the
engine buffers that differ most."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.e2e_report import run_case  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "toy_heads_flat"
os.environ["SC_ATT_GENERIC"] = "1"
a = run_case(case, "bf16", True, verbose=False)
bufs_a = {k: v.detach().clone() for k, v in a["bufs"].items()}
del os.environ["SC_ATT_GENERIC"]
b = run_case(case, "bf16", True, verbose=False)
print("generic: max_grad_rel %.3f   mma: max_grad_rel %.3f" % (a["max_grad_rel"], b["max_grad_rel"]))
rows = []
for k, va in bufs_a.items():
    vb = b["bufs"][k]
    if not va.is_floating_point():
        continue
    den = float(va.float().norm()) + 1e-20
    rows.append((float((va.float() - vb.float()).norm()) / den, k, float(va.float().abs().max())))
for r, k, mx in rows:
    if r > 0.02:
        print("%.3e  %-28s max|x|=%.3e" % (r, k, mx))
