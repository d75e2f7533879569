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
                              kv_layout=kv_layout, mae_vis_mask_ratio=cfg.get("mae_vis_mask_ratio", 0.75))
    model = SegCLIP(rh.fake_clip_state_dict(cfg), args)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.cuda().train()


def run_case(case, precision, forced=False, verbose=True):
    g = load_case(case)
    out = run_config(g["config"], g["batch"], g["param_seed"], g["batch_seed"], precision, forced, g["kv_layout"], verbose=verbose,
                     name=case)
    out["golden_loss"] = g["loss"]
    return out


def run_config(cfg, B, param_seed, batch_seed, precision, forced=False, kv="torch18_flat", verbose=True, name="", ideal=False,
               train_stem=False):
    """CUDA path vs the CPU oracle on seeded inputs.  ``ideal=True`` additionally runs the oracle with every matrix-product
    operand rounded to bf16 (oracle/bf16_emulation.py) and reports its gradient error next to the CUDA path's."""
    from segclip_b200 import _lib
    params = so.init_params(cfg, seed=param_seed)
    batch, noise = so.make_batch(cfg, B, seed=batch_seed)
    # train_stem: the parameters the reference recipe freezes are trained too (all but the fixed sin-cos decoder table)
    frozen = FROZEN_STEM[-1:] if train_stem else FROZEN_STEM
    ref_loss, ref_grads, info = so.loss_and_grads(params, batch, noise, cfg, kv, frozen=frozen)
    model = build_model(cfg, params, precision, kv)
    if train_stem:
        for n_, p_ in model.named_parameters():
            if n_ in FROZEN_STEM[:-1]:
                p_.requires_grad_(True)
    model.inject_noise({k: v.cuda() for k, v in noise.items()})
    if forced:
        f = {"main": info["assign_main"].cuda(), "pool": info["pool_arg"].cuda()}
        if cfg["use_mae"]:
            f["mae"] = info["assign_mae"].cuda()
        model.force_assignment(f)
    ids = batch["input_ids"]
    k0 = _lib.kernel_launches()
    loss = model(ids, torch.zeros_like(ids), batch["attention_mask"], batch["image"], image_seg=batch["image_seg"])
    loss.backward()
    torch.cuda.synchronize()
    k1 = _lib.kernel_launches()
    bufs = model.debug_buffers(B)
    out = dict(case=name, precision=precision, forced=forced, loss=float(loss), ref_loss=float(ref_loss),
               loss_rel=abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)),
               kernels={k: k1[k] - k0[k] for k in k1})
    idx = bufs["v.sem.idx"].cpu().long()
    out["assign_flip_rate"] = float((idx != info["assign_main"]).float().mean())
    out["pool_flip_rate"] = float((bufs["v.parg"].cpu().long() != info["pool_arg"]).float().mean())
    if cfg["use_mae"]:
        out["assign_flip_rate_mae"] = float((bufs["m.sem.idx"].cpu().long() != info["assign_mae"]).float().mean())
    errs = {}
    grads = {}
    for name_, p in model.named_parameters():
        if name_ in frozen:
            continue
        rg = ref_grads.get(name_)
        if p.grad is None:
            if rg is not None and float(rg.abs().max()) > 0:
                errs[name_] = (float("inf"), 0.0)
            continue
        mine = p.grad.detach().float().cpu()
        grads[name_] = mine
        if rg is None:
            rg = torch.zeros_like(mine)
        den = float(rg.norm()) + 1e-12
        cos = float((mine.flatten() @ rg.flatten()) / (mine.norm() * rg.norm() + 1e-30))
        errs[name_] = (float((mine - rg).norm()) / den, cos)
    out["max_grad_rel"] = max(v[0] for v in errs.values())
    out["min_grad_cos"] = min(v[1] for v in errs.values() if v[0] > 0) if errs else 1.0
    out["worst"] = sorted(((v[0], v[1], k) for k, v in errs.items()), reverse=True)[:12]
    rels = sorted(v[0] for v in errs.values())
    out["median_grad_rel"] = rels[len(rels) // 2]
    if ideal:
        from oracle.bf16_emulation import BF16Operands
        f = {"main": info["assign_main"], "pool": info["pool_arg"]}
        if cfg["use_mae"]:
            f["mae"] = info["assign_mae"]
        with BF16Operands():
            il, ig, _ = so.loss_and_grads(params, batch, noise, cfg, kv, forced=f, frozen=frozen)
        ideal_errs = {k: float((ig[k] - ref_grads[k]).norm()) / (float(ref_grads[k].norm()) + 1e-12) for k in ref_grads if k in ig}
        ir = sorted(ideal_errs.values())
        out["ideal_errs"] = ideal_errs
        out["ideal_median_grad_rel"], out["ideal_max_grad_rel"] = ir[len(ir) // 2], ir[-1]
        out["ideal_loss_rel"] = abs(float(il) - float(ref_loss)) / abs(float(ref_loss))
        # worst ratio of the CUDA path's error to the irreducible operand-rounding error (+1e-2 absolute slack)
        out["worst_vs_ideal"] = max(errs[k][0] / (ideal_errs[k] + 1e-2) for k in errs if k in ideal_errs)
    if verbose:
        print(json.dumps({k: v for k, v in out.items() if k not in ("worst", "ideal_errs")}))
        for e, c, k in out["worst"]:
            print("   %.3e  cos=%.5f  %s%s" % (e, c, k, "   (ideal %.3e)" % out["ideal_errs"].get(k, float("nan")) if ideal else ""))
    out["errs"] = errs
    out["grads"] = grads
    out["info"] = info
    out["bufs"] = bufs
    return out


if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "toy_contrastive_flat"
    precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    forced = len(sys.argv) > 3 and sys.argv[3] == "forced"
    if case.startswith("vitb16:"):        # vitb16:<batch>[:heads]  -- production-dispatch shapes, with the ideal-bf16 yardstick
        parts = case.split(":")
        heads = len(parts) > 2 and parts[2] == "heads"
        run_config(so.vit_b16_config(use_mae=heads, use_kl=heads), int(parts[1]), 5, 6, precision, forced, ideal=precision == "bf16",
                   name=case)
    else:
        run_case(case, precision, forced)
