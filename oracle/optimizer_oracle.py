"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's optimizer step for the "next" row:
torch.nn.utils.clip_grad_norm_ (main_task_align.py:326), AdaptAdamW.step (modules/optimization_adamw.py:112-174) with the
warm-up/cosine schedule (:26-30) and the logit_scale clamp (main_task_align.py:343-347).  Pinned against the reference's
own AdaptAdamW class in tests/test_optimizer.py (when /root/reference is present) and a committed golden."""
import math

import torch


def warmup_cosine(x, warmup, lr_start, lr_end):
    if x < warmup:
        return (x * (1. - lr_start) / warmup) + lr_start
    new_x = (x - warmup) / (1 - warmup)
    return lr_end + 0.5 * (1. - lr_end) * (1 + math.cos(math.pi * new_x))


def step(params, grads, state, groups, t_total, warmup, b1, b2, eps, lr_start, lr_end, clip_grad=None, clamp_max=None):
    """params/grads/state: lists of tensors / dicts (state: step, exp_avg, exp_avg_sq); groups: list of (indices, lr, wd).
    Updates params and state in place."""
    if clip_grad:
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads if g is not None)).float()
        coef = min(1.0, float(clip_grad / (total + 1e-6)))
        grads = [None if g is None else g * coef for g in grads]
    for idxs, lr, wd in groups:
        for i in idxs:
            g = grads[i]
            if g is None:
                continue
            st = state[i]
            st["step"] += 1
            t = st["step"]
            st["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
            st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (st["exp_avg_sq"].sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
            lr_t = lr * warmup_cosine(t / t_total, warmup, lr_start, lr_end) if t_total != -1 else lr
            params[i].mul_(1 - lr_t * wd)
            params[i].addcdiv_(st["exp_avg"], denom, value=-lr_t / (1 - b1 ** t))
            if clamp_max and i in clamp_max:
                params[i].clamp_(max=clamp_max[i])
