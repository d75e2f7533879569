"""Fused optimizer step: drop-in for the reference's
``clip_grad_norm_`` + ``AdaptAdamW.step()`` + ``logit_scale`` clamp (main_task_align.py:326-347,
modules/optimization_adamw.py:63-174).  Same constructor / param-group keys as ``AdaptAdamW``; ``step()`` issues two
native launches for the whole model (gradient norm, update).  No PyTorch arithmetic; fails loudly without the library."""
import ctypes as C
import math

import torch
from torch.optim import Optimizer

from . import _lib as L


def warmup_cosine(x, warmup=0.002, lr_start=0., lr_end=0.):
    """modules/optimization_adamw.py:26-30"""
    if x < warmup:
        return (x * (1. - lr_start) / warmup) + lr_start
    new_x = (x - warmup) / (1 - warmup)
    return lr_end + 0.5 * (1. - lr_end) * (1 + math.cos(math.pi * new_x))


def warmup_constant(x, warmup=0.002, lr_start=0., lr_end=0.):
    return x / warmup if x < warmup else 1.0


def warmup_linear(x, warmup=0.002, lr_start=0., lr_end=0.):
    return x / warmup if x < warmup else max((x - 1.) / (warmup - 1.), 0)


SCHEDULES = {"warmup_cosine": warmup_cosine, "warmup_constant": warmup_constant, "warmup_linear": warmup_linear}


class FusedAdaptAdamW(Optimizer):
    """Arguments as ``AdaptAdamW`` (modules/optimization_adamw.py:63-90).  ``clip_grad`` (extension, default None) fuses
    ``torch.nn.utils.clip_grad_norm_(all params, clip_grad)`` into the step; ``clamp_max`` = {param: value} fuses
    ``torch.clamp_(param.data, max=value)`` (the reference clamps ``clip.logit_scale`` to ln 100 after every step)."""

    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0, lr_start=0., lr_end=0., clip_grad=None, clamp_max=None):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if schedule not in SCHEDULES:
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                        weight_decay=weight_decay, max_grad_norm=max_grad_norm, lr_start=lr_start, lr_end=lr_end)
        super().__init__(params, defaults)
        self.clip_grad = clip_grad
        self.clamp_max = {id(p): float(v) for p, v in (clamp_max or {}).items()}
        self._table = None
        self._upload_done = None
        self._sig = None
        self._host = None

    def get_lr(self):
        lr = []
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state[p]
                if len(st) == 0:
                    return [0]
                lr.append(self._lr(group, st["step"]))
        return lr

    @staticmethod
    def _lr(group, step):
        if group["t_total"] != -1:
            return group["lr"] * SCHEDULES[group["schedule"]](step / group["t_total"], group["warmup"], group["lr_start"],
                                                             group["lr_end"])
        return group["lr"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries = []
        dev = None
        b1 = b2 = eps = None
        for group in self.param_groups:
            if b1 is None:
                b1, b2, eps = group["b1"], group["b2"], group["e"]
            elif (b1, b2, eps) != (group["b1"], group["b2"], group["e"]):
                raise L.SegclipB200Error("FusedAdaptAdamW: b1/b2/e must be equal across parameter groups")
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_cuda:
                    raise L.SegclipB200Error("FusedAdaptAdamW: fp32 CUDA parameters only (no CPU path)")
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise L.SegclipB200Error("FusedAdaptAdamW: contiguous parameters / gradients only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                t = st["step"]
                lr = self._lr(group, t)
                entries.append((p, st, 1.0 - lr * group["weight_decay"], lr / (1.0 - b1 ** t), 1.0 / math.sqrt(1.0 - b2 ** t)))
                dev = p.device
        if not entries:
            return loss
        n = len(entries)
        if self._host is None or len(self._host) != n:
            self._host = (L.OptItem * n)()
            self._sig = None
        arr = self._host
        blocks = 0
        for i, (p, st, decay, step_size, isb2) in enumerate(entries):
            it = arr[i]
            it.param, it.grad = p.data_ptr(), p.grad.data_ptr()
            it.exp_avg, it.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            it.n, it.first_block = p.numel(), blocks
            it.decay, it.step_size, it.inv_sqrt_bc2 = decay, step_size, isb2
            cm = self.clamp_max.get(id(p))
            it.clamp_max_enabled, it.clamp_max = (1, cm) if cm is not None else (0, 0.0)
            blocks += (p.numel() + 1023) // 1024
        raw = C.string_at(C.addressof(arr), C.sizeof(arr))
        if self._table is None or self._table.numel() != len(raw):
            self._table = torch.empty(len(raw), dtype=torch.uint8, device=dev)
            self._pinned = torch.empty(len(raw), dtype=torch.uint8).pin_memory()
            self._sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        if self._upload_done is not None:
            self._upload_done.synchronize()      # the previous step's async H2D copy must have read the pinned table
        self._pinned.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        self._table.copy_(self._pinned, non_blocking=True)
        self._upload_done = torch.cuda.current_stream(dev).record_event()
        lib, st_ = L.lib(), L.stream()
        max_norm = float(self.clip_grad) if self.clip_grad else 0.0
        self._sqnorm.zero_()
        if max_norm > 0:
            L.check(lib.sc_grad_sqnorm_multi(self._table.data_ptr(), n, blocks, self._sqnorm.data_ptr(), st_), "sc_grad_sqnorm_multi")
        L.check(lib.sc_adamw_multi(self._table.data_ptr(), n, blocks, self._sqnorm.data_ptr(), max_norm, b1, b2, eps, st_),
                "sc_adamw_multi")
        return loss
