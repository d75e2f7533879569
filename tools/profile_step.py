"""Per-op device-time breakdown of one training step (CUDA events around every native call).

    python tools/profile_step.py [--batch 256] [--heads] [--ncu]     (--ncu: bracket one step with
    cudaProfilerStart/Stop for `ncu --profile-from-start off`)
"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import synthetic_batch                      # noqa: E402
from segclip_b200 import config                        # noqa: E402
from segclip_b200 import _lib as L                     # noqa: E402
from segclip_b200.modeling import SegCLIP              # noqa: E402


def describe(op):
    if op.name == "sc_gemm":
        d = op.keep[0]
        kind = "tc" if d.in_dtype == L.BF16 else "simt"
        return "gemm_%s %s%s M=%d N=%d K=%d act=%d res=%d c2=%d acc=%d" % (
            kind, "T" if d.trans_a else "N", "T" if d.trans_b else "N", d.M, d.N, d.K, d.act, int(bool(d.residual)),
            int(bool(d.C2)), d.accumulate)
    if op.name in ("sc_attention_fwd", "sc_attention_bwd"):
        a = op.keep[0] if op.name == "sc_attention_fwd" else op.keep[0].fwd
        return "%s B=%d H=%d Lq=%d Lk=%d hd=%d" % (op.name, a.B, a.H, a.Lq, a.Lk, a.hd)
    if op.name in ("sc_layernorm_fwd", "sc_layernorm_bwd"):
        d = op.keep[0]
        return "%s rows=%d D=%d" % (op.name, d.rows, d.D)
    return op.name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--heads", action="store_true")
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--model", default="vitb16", choices=["vitb16", "vitl14"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = (config.vit_l14 if args.model == "vitl14" else config.vit_b16)(use_mae=args.heads, use_kl=args.heads)
    tc = argparse.Namespace(local_rank=0, rank=0, world_size=1, first_stage_layer=10, use_vision_mae_recon=args.heads,
                            use_seglabel=args.heads, precision="bf16")
    torch.manual_seed(0)
    model = SegCLIP(config.shape_state_dict(cfg), tc).to(dev).train()
    batch = {k: v.to(dev) for k, v in synthetic_batch(cfg, args.batch, 0, 0, args.heads).items()}

    def step():
        model.zero_grad(set_to_none=True)
        loss = model(batch["input_ids"], None, None, batch["image"], image_seg=batch["image_seg"] if args.heads else None)
        loss.backward()

    step()
    step()
    torch.cuda.synchronize()
    if args.ncu:
        torch.cuda.cudart().cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    eng = model._engine
    pl = eng.plan(args.batch)
    st = L.stream()
    recs = []
    torch._foreach_zero_(pl.zero)
    eng.gflat.zero_()
    for phase, tape in (("fwd", pl.fwd), ("bwd", pl.bwd)):
        for op in tape:
            if isinstance(op, str):
                continue
            if isinstance(op, tuple):
                op = op[0]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op(st)
            e1.record()
            recs.append((phase, describe(op), e0, e1))
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    tot = 0.0
    for phase, name, e0, e1 in recs:
        ms = e0.elapsed_time(e1)
        tot += ms
        k = (phase, name)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + ms)
    print("total %.2f ms over %d native calls (%s, batch %d, heads=%s)" % (tot, len(recs), args.model, args.batch, args.heads))
    byname = collections.Counter()
    for (phase, name), (c, t) in agg.items():
        byname[name.split(" ")[0] + ":" + phase] += t
    print("--- by kernel family")
    for k, t in byname.most_common():
        print("  %-28s %9.3f ms  %5.1f%%" % (k, t, 100 * t / tot))
    print("--- top entries")
    for (phase, name), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print("  %s x%-3d %9.3f ms  %5.1f%%  %s" % (phase, c, t, 100 * t / tot, name))


if __name__ == "__main__":
    main()
