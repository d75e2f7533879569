"""BASELINE configs[4]: block-kernel roofline sweep -- ONE pre-LN residual attention block (module_seg_vit.py:162-196),
forward + backward, at ViT-L width (1024, 16 heads) for seq_len 196 and 577, through the same native ops the engine
replays.  Prints per-kernel device times and the block's achieved TFLOP/s against its algorithmic FLOPs
(24 L D^2 + 4 L^2 D per sample forward, x3 forward+backward).

    python tools/block_sweep.py [--batch 64] [--width 1024] [--seq 196 577]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segclip_b200 import ops  # noqa: E402


def build_block(B, L, D, H):
    dev, bf, f32 = "cuda", torch.bfloat16, torch.float32
    M = B * L
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, dt=f32, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(dt)
    W = dict(qkv=rn(3 * D, D, dt=bf, sc=D ** -0.5), o=rn(D, D, dt=bf, sc=0.02), fc=rn(4 * D, D, dt=bf, sc=0.02), pr=rn(D, 4 * D, dt=bf, sc=0.02))
    bias = dict(qkv=rn(3 * D, sc=0.02), o=rn(D, sc=0.02), fc=rn(4 * D, sc=0.02), pr=rn(D, sc=0.02))
    gW = {k: torch.zeros_like(v, dtype=f32) for k, v in W.items()}
    gB = {k: torch.zeros_like(v) for k, v in bias.items()}
    ln = {k: (torch.ones(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)) for k in ("1", "2")}
    x = rn(M, D)
    h1, h2 = torch.empty(M, D, device=dev, dtype=bf), torch.empty(M, D, device=dev, dtype=bf)
    st1 = (torch.empty(M, device=dev), torch.empty(M, device=dev))
    st2 = (torch.empty(M, device=dev), torch.empty(M, device=dev))
    qkv, att = torch.empty(M, 3 * D, device=dev, dtype=bf), torch.empty(M, D, device=dev, dtype=bf)
    lse = torch.empty(B, H, L, device=dev)
    x_mid, x_out = torch.empty(M, D, device=dev), torch.empty(M, D, device=dev)
    hpre, hact = torch.empty(M, 4 * D, device=dev, dtype=bf), torch.empty(M, 4 * D, device=dev, dtype=bf)
    s3 = (L * 3 * D, 3 * D)
    ad = ops.attn_desc(qkv, qkv[:, D:], qkv[:, 2 * D:], att, lse, B, H, L, L, D // H, s3, s3, s3, (L * D, D))
    fwd = [("ln1", ops.layernorm_op(x, ln["1"][0], ln["1"][1], h1, 1e-5, *st1)),
           ("qkv", ops.gemm_op(h1, W["qkv"], qkv, bias=bias["qkv"])),
           ("attention", ops.attention_op(ad)),
           ("out_proj", ops.gemm_op(att, W["o"], x_mid, bias=bias["o"], residual=x)),
           ("ln2", ops.layernorm_op(x_mid, ln["2"][0], ln["2"][1], h2, 1e-5, *st2)),
           ("c_fc", ops.gemm_op(h2, W["fc"], hact, bias=bias["fc"], act=ops.ACT_QUICKGELU, C2=hpre)),
           ("c_proj", ops.gemm_op(hact, W["pr"], x_out, bias=bias["pr"], residual=x_mid))]
    dx = rn(M, D, dt=bf)
    d_a, d_ln, d_att, dqkv = (torch.empty(M, 4 * D, device=dev, dtype=bf), torch.empty(M, D, device=dev, dtype=bf),
                              torch.empty(M, D, device=dev, dtype=bf), torch.empty(M, 3 * D, device=dev, dtype=bf))
    delta = torch.empty(B, H, L, device=dev)
    bwd = [("c_proj dgrad*", ops.gemm_op(dx, W["pr"], d_a, trans_b=True, mul_aux=hpre, mul_aux_act=ops.ACT_QUICKGELU, colsum_out=gB["fc"])),
           ("c_proj wgrad", ops.gemm_op(dx, hact, gW["pr"], trans_a=True, trans_b=True, accumulate=True, split_k=-1)),
           ("c_fc dgrad", ops.gemm_op(d_a, W["fc"], d_ln, trans_b=True)),
           ("c_fc wgrad", ops.gemm_op(d_a, h2, gW["fc"], trans_a=True, trans_b=True, accumulate=True, split_k=-1)),
           ("ln2 bwd", ops.layernorm_bwd_op(d_ln, x_mid, st2[0], st2[1], ln["2"][0], dx, True, None, ln["2"][2], ln["2"][3], dx_colsum=gB["o"])),
           ("out_proj dgrad", ops.gemm_op(dx, W["o"], d_att, trans_b=True)),
           ("out_proj wgrad", ops.gemm_op(dx, att, gW["o"], trans_a=True, trans_b=True, accumulate=True, split_k=-1)),
           ("attention bwd", ops.attention_bwd_op(ad, d_att, dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta)),
           ("qkv dgrad", ops.gemm_op(dqkv, W["qkv"], d_ln, trans_b=True)),
           ("qkv wgrad", ops.gemm_op(dqkv, h1, gW["qkv"], trans_a=True, trans_b=True, accumulate=True, split_k=-1)),
           ("qkv bias", ops.colsum_op(dqkv, gB["qkv"])),
           ("ln1 bwd", ops.layernorm_bwd_op(d_ln, x, st1[0], st1[1], ln["1"][0], dx, True, None, ln["1"][2], ln["1"][3], dx_colsum=gB["pr"]))]
    return fwd, bwd, (x, W, bias, gW, gB, ln)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--seq", type=int, nargs="+", default=[196, 577])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    D, H = args.width, args.width // 64
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    for L in args.seq:
        B = args.batch
        fwd, bwd, keep = build_block(B, L, D, H)
        for _ in range(2):
            for _, op in fwd + bwd:
                op()
        torch.cuda.synchronize()
        times = {}
        for name, op in fwd + bwd:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                op()
            e1.record()
            torch.cuda.synchronize()
            times[name] = e0.elapsed_time(e1) / args.reps
        f_fwd = B * (24.0 * L * D * D + 4.0 * L * L * D)
        t_fwd = sum(times[n] for n, _ in fwd)
        t_all = sum(times.values())
        print("# one block, width %d, %d heads, batch %d, seq_len %d  (algorithmic %.1f GF fwd / sample)" % (D, H, B, L, f_fwd / B / 1e9))
        for n, t in times.items():
            print("  %-16s %8.3f ms" % (n, t))
        print(json.dumps({"seq_len": L, "width": D, "batch": B, "ms_fwd": t_fwd, "ms_fwd_bwd": t_all,
                          "tflops_fwd": f_fwd / t_fwd / 1e9, "tflops_fwd_bwd": 3 * f_fwd / t_all / 1e9,
                          "frac_fwd": f_fwd / t_fwd / 1e9 / peak, "frac_fwd_bwd": 3 * f_fwd / t_all / 1e9 / peak, "peak": peak,
                          "attention_share": (times["attention"] + times["attention bwd"]) / t_all}))
        del fwd, bwd, keep
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
