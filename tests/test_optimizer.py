"""Fused optimizer step.  CPU: the oracle restatement against the reference's AdaptAdamW class
(imported from /root/reference when present).  GPU: FusedAdaptAdamW against the oracle over several steps."""
import math
import sys

import pytest
import torch

from oracle import optimizer_oracle as oo
from oracle import ref_harness as rh

SHAPES = [(37, 19), (1000,), (), (3, 5, 7), (1025,)]
HP = dict(t_total=50, warmup=0.1, b1=0.9, b2=0.98, eps=1e-6, lr_start=0.0, lr_end=0.0)
GROUPS = [([0, 2], 4e-3, 0.2), ([1, 3], 4e-6, 0.0), ([4], 4e-3, 0.2)]


def _make(seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(s, generator=g).to(device) for s in SHAPES]
    grads_per_step = [[torch.randn(s, generator=g).to(device) * (3.0 if k == 0 else 0.3) for s in SHAPES] for k in range(4)]
    return params, grads_per_step


def _run_oracle(params, grads_per_step, clip, clamp):
    params = [p.clone() for p in params]
    state = [dict(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)) for p in params]
    for grads in grads_per_step:
        oo.step(params, grads, state, GROUPS, clip_grad=clip, clamp_max=clamp, **HP)
    return params


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")
def test_oracle_matches_reference_adaptadamw():
    rh._install_shims()
    from modules.optimization_adamw import AdaptAdamW            # the reference's optimizer, unmodified
    params, gps = _make(0)
    ref_p = [torch.nn.Parameter(p.clone()) for p in params]
    opt = AdaptAdamW([dict(params=[ref_p[i] for i in idx], lr=lr, weight_decay=wd) for idx, lr, wd in GROUPS], lr=4e-3,
                     warmup=HP["warmup"], schedule="warmup_cosine", b1=HP["b1"], b2=HP["b2"], e=HP["eps"],
                     t_total=HP["t_total"], weight_decay=0.2, max_grad_norm=1.0, lr_start=0.0, lr_end=0.0)
    for grads in gps:
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref_p, 1.0)              # main_task_align.py:326
        opt.step()
        opt.zero_grad()
        torch.clamp_(ref_p[2].data, max=math.log(100))          # main_task_align.py:343-347 (logit_scale is a scalar)
    mine = _run_oracle(params, gps, clip=1.0, clamp={2: math.log(100)})
    for a, b in zip(mine, ref_p):
        assert torch.allclose(a, b.data, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("clip", [None, 1.0])
def test_fused_step_matches_oracle(clip):
    from segclip_b200.optim import FusedAdaptAdamW
    params, gps = _make(1)
    want = _run_oracle(params, gps, clip=clip, clamp={2: 0.05})
    dev_p = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = FusedAdaptAdamW([dict(params=[dev_p[i] for i in idx], lr=lr, weight_decay=wd) for idx, lr, wd in GROUPS], lr=4e-3,
                          warmup=HP["warmup"], schedule="warmup_cosine", b1=HP["b1"], b2=HP["b2"], e=HP["eps"],
                          t_total=HP["t_total"], weight_decay=0.2, lr_start=0.0, lr_end=0.0, clip_grad=clip,
                          clamp_max={dev_p[2]: 0.05})
    for grads in gps:
        for p, g in zip(dev_p, grads):
            p.grad = g.clone().cuda()
        opt.step()
        opt.zero_grad()
    torch.cuda.synchronize()
    for a, b in zip(dev_p, want):
        assert torch.allclose(a.data.cpu(), b, rtol=2e-5, atol=1e-6), float((a.data.cpu() - b).abs().max())


def test_fused_optimizer_refuses_cpu_tensors():
    from segclip_b200 import _lib
    from segclip_b200.optim import FusedAdaptAdamW
    p = torch.nn.Parameter(torch.randn(4))
    p.grad = torch.randn(4)
    with pytest.raises(_lib.SegclipB200Error):
        FusedAdaptAdamW([p], lr=1e-3).step()
