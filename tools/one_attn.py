"""Runs one self-attention fwd+bwd (for ncu captures / A-B timing).  usage: one_attn.py [vision|text]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "vision"
B, H, L, hd, causal = (256, 12, 196, 64, False) if which == "vision" else (256, 8, 77, 64, True)
D = H * hd
qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
do = torch.randn(B * L, D, device="cuda").bfloat16()
o = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, L, device="cuda")
st = (L * 3 * D, 3 * D)
a = ops.attn_desc(qkv, qkv[:, D:], qkv[:, 2 * D:], o, lse, B, H, L, L, hd, st, st, st, (L * D, D), causal)
dqkv = torch.empty_like(qkv)
delta = torch.empty(B, H, L, device="cuda")
f, b = ops.attention_op(a), ops.attention_bwd_op(a, do, dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta)
for _ in range(3):
    f()
    b()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); f(); e[1].record(); b(); e[2].record()
torch.cuda.synchronize()
print("%s: fwd %.3f ms  bwd %.3f ms" % (which, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
