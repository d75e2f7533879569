"""oracle/segclip_oracle.py against the committed golden vectors (made by the unmodified
reference, tests/golden/make_golden.py).  Pure CPU."""
import pytest
import torch

from oracle import segclip_oracle as so
from tests.golden_util import compare_grads, load_case

FROZEN = ("vis_mae_decoder.decoder_pos_embed",)
CASES = ["toy_contrastive_flat", "toy_heads_flat", "toy_heads_per_sample",
         "vitb16_contrastive_b2", "vitb16_heads_b2"]


@pytest.mark.parametrize("case", CASES)
def test_port_matches_reference_golden(case):
    g = load_case(case)
    cfg = g["config"]
    params = so.init_params(cfg, seed=g["param_seed"])
    batch, noise = so.make_batch(cfg, g["batch"], seed=g["batch_seed"])
    loss, grads, _ = so.loss_and_grads(params, batch, noise, cfg, g["kv_layout"], frozen=FROZEN)
    assert abs(float(loss) - g["loss"]) <= 1e-5 * abs(g["loss"])
    assert not compare_grads(grads, g["grads"], tol=5e-4)


@pytest.mark.parametrize("world", [2, 8])
def test_port_multi_rank_matches_reference_golden(world):
    """W=2 / W=8: diffdist all-gather + DDP mean of the reference (W gloo processes when the fixture was
    made) vs the single-process multi-rank formulation of the port."""
    g = load_case("toy_heads_flat_w%d" % world)
    cfg = g["config"]
    p = {k: v.clone().requires_grad_(v.is_floating_point() and k not in FROZEN)
         for k, v in so.init_params(cfg, seed=g["param_seed"]).items()}
    bn = [so.make_batch(cfg, g["batch"], seed=g["batch_seed"], rank=r) for r in range(world)]
    losses, _ = so.forward_multi_rank(p, [b for b, _ in bn], [n for _, n in bn], cfg)
    for mine, ref in zip(losses, g["loss"]):
        assert abs(float(mine.detach()) - ref) <= 1e-5 * abs(ref)
    (sum(losses) / world).backward()
    grads = {k: v.grad for k, v in p.items() if v.grad is not None}
    assert not compare_grads(grads, g["grads"], tol=5e-4)
