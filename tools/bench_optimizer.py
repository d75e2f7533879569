"""Fused optimizer step (clip + AdaptAdamW + clamp) on the real ViT-B/16 parameter set: time and achieved HBM GB/s
(28 B per parameter: read p, g, m, v; write p, m, v), next to the reference's per-tensor PyTorch loop on the same GPU."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_harness import fake_clip_state_dict   # noqa: E402
from oracle import segclip_oracle as so                # noqa: E402
from oracle import optimizer_oracle as oo              # noqa: E402
from segclip_b200.modeling import SegCLIP              # noqa: E402
from segclip_b200.optim import FusedAdaptAdamW         # noqa: E402

cfg = so.vit_b16_config(use_mae=True, use_kl=True)
tc = argparse.Namespace(first_stage_layer=10, use_vision_mae_recon=True, use_seglabel=True)
model = SegCLIP(fake_clip_state_dict(cfg), tc).cuda()
params = [p for p in model.parameters() if p.requires_grad]
n = sum(p.numel() for p in params)
for p in params:
    p.grad = torch.randn_like(p) * 0.01
opt = FusedAdaptAdamW(params, lr=4e-3, warmup=0.1, t_total=1000, schedule="warmup_cosine", b1=0.9, b2=0.98, e=1e-6,
                      weight_decay=0.2, clip_grad=1.0, clamp_max={model.clip.logit_scale: 4.6052})
for _ in range(3):
    opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(10):
    opt.step()
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10 * 1e3
ms = e0.elapsed_time(e1) / 10
print("fused clip+AdaptAdamW+clamp: %d tensors, %.1f M params: %.3f ms device (%.3f ms wall), %.0f GB/s (28 B/param + 4 B/param norm pass)"
      % (len(params), n / 1e6, ms, wall, n * 32 / ms / 1e6))
# reference-style per-tensor loop (oracle restatement executed with torch ops on the GPU)
st = [dict(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)) for p in params]
data = [p.data for p in params]
grads = [p.grad for p in params]
groups = [(list(range(len(params))), 4e-3, 0.2)]
for _ in range(2):
    oo.step(data, grads, st, groups, 1000, 0.1, 0.9, 0.98, 1e-6, 0.0, 0.0, clip_grad=1.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    oo.step(data, grads, st, groups, 1000, 0.1, 0.9, 0.98, 1e-6, 0.0, 0.0, clip_grad=1.0)
torch.cuda.synchronize()
print("reference-style per-tensor PyTorch loop on the same GPU: %.3f ms wall per step" % ((time.perf_counter() - t0) / 5 * 1e3))
