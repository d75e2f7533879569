"""Debug report: CUDA path vs the CPU oracle on one golden case (loss, assignment, per-parameter
gradient errors).  Test infrastructure -- imports oracle/.

    python tools/e2e_report.py toy_heads_flat fp32
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402  (only fake_clip_state_dict: shapes)
from oracle import segclip_oracle as so       # noqa: E402
from segclip_b200.engine import FROZEN_STEM   # noqa: E402
from segclip_b200.modeling import SegCLIP     # noqa: E402
from tests.golden_util import load_case       # noqa: E402


def build_model(cfg, params, precision, kv_layout):
    args = argparse.Namespace(local_rank=0, rank=0, world_size=1, first_stage_layer=cfg["first_stage_layer"],
                              use_vision_mae_recon=cfg["use_mae"], use_seglabel=cfg["use_kl"], precision=precision,
                              kv_layout=kv_layout)
    model = SegCLIP(rh.fake_clip_state_dict(cfg), args)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.cuda().train()


def run_case(case, precision, forced=False, verbose=True):
    g = load_case(case)
    cfg, kv = g["config"], g["kv_layout"]
    params = so.init_params(cfg, seed=g["param_seed"])
    batch, noise = so.make_batch(cfg, g["batch"], seed=g["batch_seed"])
    ref_loss, ref_grads, info = so.loss_and_grads(params, batch, noise, cfg, kv, frozen=FROZEN_STEM)
    model = build_model(cfg, params, precision, kv)
    model.inject_noise({k: v.cuda() for k, v in noise.items()})
    if forced:
        f = {"main": info["assign_main"].cuda(), "pool": info["pool_arg"].cuda()}
        if cfg["use_mae"]:
            f["mae"] = info["assign_mae"].cuda()
        model.force_assignment(f)
    ids = batch["input_ids"]
    loss = model(ids, torch.zeros_like(ids), batch["attention_mask"], batch["image"], image_seg=batch["image_seg"])
    loss.backward()
    torch.cuda.synchronize()
    bufs = model.debug_buffers(g["batch"])
    out = dict(case=case, precision=precision, forced=forced, loss=float(loss), ref_loss=float(ref_loss),
               golden_loss=g["loss"], loss_rel=abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)))
    idx = bufs["v.sem.idx"].cpu().long()
    out["assign_flip_rate"] = float((idx != info["assign_main"]).float().mean())
    out["pool_flip_rate"] = float((bufs["v.parg"].cpu().long() != info["pool_arg"]).float().mean())
    if cfg["use_mae"]:
        out["assign_flip_rate_mae"] = float((bufs["m.sem.idx"].cpu().long() != info["assign_mae"]).float().mean())
    errs = {}
    for name, p in model.named_parameters():
        if name in FROZEN_STEM:
            continue
        rg = ref_grads.get(name)
        if p.grad is None:
            if rg is not None and float(rg.abs().max()) > 0:
                errs[name] = (float("inf"), 0.0)
            continue
        mine = p.grad.detach().float().cpu()
        if rg is None:
            rg = torch.zeros_like(mine)
        den = float(rg.norm()) + 1e-12
        cos = float((mine.flatten() @ rg.flatten()) / (mine.norm() * rg.norm() + 1e-30))
        errs[name] = (float((mine - rg).norm()) / den, cos)
    out["max_grad_rel"] = max(v[0] for v in errs.values())
    out["min_grad_cos"] = min(v[1] for v in errs.values() if v[0] > 0) if errs else 1.0
    out["worst"] = sorted(((v[0], v[1], k) for k, v in errs.items()), reverse=True)[:12]
    if verbose:
        print(json.dumps({k: v for k, v in out.items() if k != "worst"}))
        for e, c, k in out["worst"]:
            print("   %.3e  cos=%.5f  %s" % (e, c, k))
    out["errs"] = errs
    out["info"] = info
    out["bufs"] = bufs
    return out


if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "toy_contrastive_flat"
    precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    forced = len(sys.argv) > 3 and sys.argv[3] == "forced"
    run_case(case, precision, forced)
