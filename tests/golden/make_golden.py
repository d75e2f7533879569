"""Generates tests/golden/*.json by running the UNMODIFIED reference (/root/reference) on CPU.

    python tests/golden/make_golden.py            # all cases (about a minute)

Each case fixes: config, parameter seed (oracle.segclip_oracle.init_params), batch seed
(make_batch), kv layout, world size.  The fixture stores the reference's loss and, for every
parameter that received a gradient, its L2 norm and a checksum <grad, r_name> against a seeded
random vector (so element permutations are detected without committing full tensors).
"""
import hashlib
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as rh        # noqa: E402
from oracle import segclip_oracle as so     # noqa: E402


def checksum_vector(name, numel):
    seed = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
    return torch.randn(numel, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


def summarize(grads):
    out = {}
    for name, g in grads.items():
        g64 = g.double().flatten()
        out[name] = [float(g64.norm()), float(g64 @ checksum_vector(name, g64.numel()))]
    return out


CASES = {
    # name: (config factory, overrides, B, kv_layout, param seed, batch seed)
    "toy_contrastive_flat": ("toy", dict(), 3, "torch18_flat", 1, 2),
    "toy_heads_flat": ("toy", dict(use_mae=True, use_kl=True), 3, "torch18_flat", 1, 2),
    "toy_heads_per_sample": ("toy", dict(use_mae=True, use_kl=True), 5, "per_sample", 3, 4),
    "vitb16_contrastive_b2": ("vitb16", dict(), 2, "torch18_flat", 0, 0),
    "vitb16_heads_b2": ("vitb16", dict(use_mae=True, use_kl=True), 2, "torch18_flat", 0, 0),
}


def make_cfg(kind, over):
    return so.toy_config(**over) if kind == "toy" else so.vit_b16_config(**over)


def run_case(name):
    kind, over, B, kv, pseed, bseed = CASES[name]
    cfg = make_cfg(kind, over)
    model = rh.build_reference_model(cfg, kv_layout=kv)
    params = so.init_params(cfg, seed=pseed)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    batch, noise = so.make_batch(cfg, B, seed=bseed)
    loss, grads = rh.run_reference(model, batch, noise, cfg["use_mae"], kv)
    return dict(case=name, config=cfg, batch=B, kv_layout=kv, param_seed=pseed, batch_seed=bseed,
                world=1, loss=float(loss), grads=summarize(grads), torch=torch.__version__)


def _w2_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2 if world <= 2 else 1)
    cfg = so.toy_config(use_mae=True, use_kl=True)
    rh.init_dist(rank, world, port)
    model = rh.build_reference_model(cfg, rank=rank, world=world)
    params = so.init_params(cfg, seed=5)
    model.load_state_dict(params, strict=False)
    batch, noise = so.make_batch(cfg, 3, seed=6, rank=rank)
    loss, grads = rh.run_reference(model, batch, noise, True)
    for g in grads.values():                      # what DDP leaves in .grad: the rank mean
        dist.all_reduce(g)
        g /= world
    if rank == 0:
        losses = [torch.zeros(()) for _ in range(world)]
    dist.gather(loss, losses if rank == 0 else None, dst=0)
    if rank == 0:
        q.put(dict(case="toy_heads_flat_w%d" % world, config=cfg, batch=3, kv_layout="torch18_flat", param_seed=5,
                   batch_seed=6, world=world, loss=[float(x) for x in losses], grads=summarize(grads),
                   torch=torch.__version__))
    dist.barrier()


def run_w2(world=2):
    """W gloo processes of the unmodified reference (diffdist all-gather + DDP-style gradient mean)."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_w2_worker, args=(r, world, 29541 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join()
    return out


if __name__ == "__main__":
    assert rh.reference_available(), "needs /root/reference"
    torch.manual_seed(0)
    results = [run_case(n) for n in CASES]
    results.append(run_w2(2))
    results.append(run_w2(8))       # full fan-out of the 8-GPU box (epoch / consumed protocol of the P2P exchange)
    for r in results:
        path = os.path.join(HERE, r["case"] + ".json")
        with open(path, "w") as f:
            json.dump(r, f, indent=0, sort_keys=True)
        print(r["case"], r["loss"], len(r["grads"]), "grads ->", os.path.relpath(path))
