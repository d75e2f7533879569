"""Debug: cycle-stamped timeline of CTA 0 of the tcgen05 attention kernels (library built with -DSC_ATT_TRACE).

    SEGCLIP_B200_LIB=<trace build>/libsegclip_b200.so python tools/trace_attn.py [vision|text] [fwd|bwd]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import ops, _lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "vision"
phase = sys.argv[2] if len(sys.argv) > 2 else "bwd"
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
f(); b()
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 2048)()
n = (ctypes.c_int * 2)()
lib.sc_debug_attn_trace(buf, n)          # drop the warm-up trace
(b if phase == "bwd" else f)()
lib.sc_debug_attn_trace(buf, n)
names = {1: "item start", 2: "loads landed", 3: "SdP(0) issued", 4: "bar_p seen", 5: "next SdP issued / acc free", 6: "dVdKdQ issued",
         7: "last MMAs retired", 10: "soft: wait S", 11: "soft: S ready", 12: "soft: computed", 13: "soft: tiles free", 14: "soft: P arrived",
         15: "soft: dkv ready", 16: "soft: dkv stored", 17: "soft: dq stored", 18: "soft: S/dP loaded from TMEM",
         20: "F item start", 21: "F QK landed", 22: "F S issued", 23: "F S done", 24: "F V+P ready", 25: "F PV issued", 26: "F O done",
         27: "F O read out", 30: "Fs wait S", 31: "Fs S ready", 32: "Fs pass1 done", 33: "Fs P arrived", 34: "Fs O ready",
         35: "Fs O loaded", 36: "Fs stored"}
print("# %s attention %s, B=%d H=%d L=%d" % (which, phase, B, H, L))
for who in range(2):
    print("== %s  (%d events)" % (("control warp", "softmax warp 0")[who], n[who]))
    t0 = None
    prev = None
    for i in range(min(n[who], 150)):
        tag, clk = buf[who * 512 + 2 * i], buf[who * 512 + 2 * i + 1]
        t0 = clk if t0 is None else t0
        print("  %8d  (+%6d)  %s" % (clk - t0, 0 if prev is None else clk - prev, names.get(tag, tag)))
        prev = clk
