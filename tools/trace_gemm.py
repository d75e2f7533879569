"""Debug: start / end time of every CTA of one 2-CTA GEMM launch (library built with -DSC_GEMM_TRACE): the spread of the
finish times is what a dynamic tile scheduler could recover from the static round-robin assignment.

    SC_LIB_DIR=/tmp/trace SC_NVCC_EXTRA=-DSC_GEMM_TRACE python -m segclip_b200.build
    SEGCLIP_B200_LIB=/tmp/trace/libsegclip_b200.so python tools/trace_gemm.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import _lib, ops  # noqa: E402

dev = "cuda"
lib = ctypes.CDLL(_lib.LIB_PATH)
SHAPES = [("qkv fwd", 50176, 2304, 768, {}), ("c_proj fwd", 50176, 768, 3072, {}), ("c_fc dgrad", 50176, 768, 3072, {"tb": True}),
          ("text qkv", 19712, 1536, 512, {}), ("plain 8192^3", 8192, 8192, 8192, {})]
for name, M, N, K, kw in SHAPES:
    A = torch.randn(M, K, device=dev).bfloat16()
    B = (torch.randn(K, N, device=dev) if kw.get("tb") else torch.randn(N, K, device=dev)).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    op = ops.gemm_op(A, B, out, trans_b=bool(kw.get("tb")))
    for rep in range(4):
        op()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 512)()
    lib.sc_debug_gemm_trace(buf)
    st, en = list(buf[:148]), list(buf[256:256 + 148])
    t0 = min(st)
    ends = sorted(e - t0 for e in en)
    dur = ends[-1]
    print("%-14s M=%d N=%d K=%d: kernel %.1f us; CTA start spread %.1f us; CTA finish: first %.1f  median %.1f  last %.1f us  "
          "(idle tail of the median CTA %.1f %%)" % (name, M, N, K, dur / 1e3, (max(st) - t0) / 1e3, ends[0] / 1e3, ends[74] / 1e3,
                                                   ends[-1] / 1e3, 100.0 * (ends[-1] - ends[74]) / dur))
    # by SM-pair position: which pairs are slow
    pairs = sorted(((en[2 * i] - t0) / 1e3, i) for i in range(74))
    print("   slowest pairs:", ", ".join("%d:%.1f" % (i, t) for t, i in pairs[-6:]), "  fastest:", ", ".join("%d:%.1f" % (i, t) for t, i in pairs[:4]))
