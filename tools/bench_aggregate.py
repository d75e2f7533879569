"""Isolated timing of the fused aggregation kernels (sc_assign_aggregate_fwd / sc_assign_bwd) at the benchmark shape.

    python tools/bench_aggregate.py [B L D]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segclip_b200 import ops  # noqa: E402


def main():
    B, L, D = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 196, 768)
    G, dev = 8, "cuda"
    torch.manual_seed(0)
    qf = torch.randn(B * G, D, device=dev) / D ** 0.25
    k = torch.randn(B * L, D, device=dev) / D ** 0.25
    v = torch.randn(B * L, D, device=dev).bfloat16()
    u = torch.rand(B, G, L, device=dev)
    y, soft = torch.empty(B, G, L, device=dev), torch.empty(B, G, L, device=dev)
    idx, count = torch.empty(B, L, device=dev, dtype=torch.int32), torch.zeros(B, G, device=dev)
    agg, ssum = torch.empty(B * G, D, device=dev), torch.empty(B * G, D, device=dev)
    fwd = ops.assign_aggregate_fwd_op(qf, k, u, y, idx, count, v, agg, ssum, B, L, D, 0.9, None, soft)
    d_logits, d_v, d_k = torch.empty(B, G, L, device=dev), torch.empty_like(v), torch.empty_like(k)
    d_qf, dsum = torch.empty_like(qf), torch.randn(B * G, D, device=dev)
    bwd = ops.assign_bwd_op(dsum, agg, v, idx, count, y, None, qf, k, d_logits, d_v, d_k, dsum, d_qf, B, L, D, 0.9)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, op, nbytes in (("assign_aggregate_fwd", fwd, k.numel() * 4 + v.numel() * 2),
                             ("assign_bwd", bwd, k.numel() * 8 + v.numel() * 4)):
        op()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print("%-22s B=%d L=%d D=%d  %.3f ms  %.0f GB/s (algorithmic bytes %.0f MB)" % (name, B, L, D, ms, nbytes / ms / 1e6, nbytes / 1e6))


if __name__ == "__main__":
    main()
