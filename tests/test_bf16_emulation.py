"""The "ideal bf16" yardstick (oracle/bf16_emulation.py) behaves as documented: it rounds the operands of every matrix
product (forward and backward), leaves the plain oracle untouched outside the context, and stays within a few percent of
the fp32 gradients.  Pure CPU."""
import torch

from oracle import segclip_oracle as so
from oracle.bf16_emulation import BF16Operands

FROZEN = ("vis_mae_decoder.decoder_pos_embed",)


def test_operand_rounding_mode():
    cfg = so.toy_config(use_mae=True, use_kl=True)
    params = so.init_params(cfg, seed=1)
    batch, noise = so.make_batch(cfg, 3, seed=2)
    l0, g0, i0 = so.loss_and_grads(params, batch, noise, cfg, frozen=FROZEN)
    forced = {"main": i0["assign_main"], "mae": i0["assign_mae"], "pool": i0["pool_arg"]}
    with BF16Operands() as mode:
        l1, g1, i1 = so.loss_and_grads(params, batch, noise, cfg, forced=forced, frozen=FROZEN)
    assert mode.products > 100                                          # every Linear / matmul / einsum / conv went through it
    assert torch.equal(i1["pool_arg"], i0["pool_arg"]) and torch.equal(i1["assign_main"], i0["assign_main"])
    l2, g2, _ = so.loss_and_grads(params, batch, noise, cfg, frozen=FROZEN)
    assert float(l2) == float(l0) and all(torch.equal(g2[k], g0[k]) for k in g0)     # no leak outside the context
    assert 0 < abs(float(l1) - float(l0)) <= 1e-2 * abs(float(l0))
    rels = sorted(float((g1[k] - g0[k]).norm() / (g0[k].norm() + 1e-12)) for k in g0)
    assert 1e-3 < rels[len(rels) // 2] < 0.1 and rels[-1] < 0.3, (rels[len(rels) // 2], rels[-1])


def test_rounding_function_values_and_gradients():
    x = torch.tensor([1.0 + 2 ** -10, 3.0], requires_grad=True)
    w = torch.tensor([[1.0, 1.0]])
    with BF16Operands():
        y = torch.nn.functional.linear(x.unsqueeze(0), w)
    assert float(y) == 4.0                                               # 1 + 2^-10 rounds to 1 in bf16
    y.backward(torch.tensor([[1.0 + 2 ** -10]]))
    assert torch.equal(x.grad, torch.tensor([1.0, 1.0]))                 # the incoming gradient is rounded too
