"""Where the end-to-end number loses against the device-resident one: the same step timed with
(resident | pinned host) inputs x (no loss read | float(loss) every step).

    python tools/e2e_gap.py [--batch 256] [--steps 10]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--heads", action="store_true")
    ap.add_argument("--model", default="vitb16")
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from segclip_b200.config import shape_state_dict
    from segclip_b200.modeling import SegCLIP
    cfg = bench.model_config(a)
    dev = torch.device("cuda", 0)
    tc = argparse.Namespace(local_rank=0, rank=0, world_size=1, first_stage_layer=10, use_vision_mae_recon=a.heads,
                            use_seglabel=a.heads, precision=a.precision)
    torch.manual_seed(0)
    model = SegCLIP(shape_state_dict(cfg), tc).to(dev).train()
    host = bench.synthetic_batch(cfg, a.batch, 0, 0, a.heads)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    u8 = dict(pinned)
    u8["image"] = (torch.rand_like(host["image"]) * 255).to(torch.uint8).pin_memory()

    def step(src, read):
        model.zero_grad(set_to_none=True)
        loss = model(src["input_ids"], None, None, src["image"], image_seg=src["image_seg"] if a.heads else None)
        loss.backward()
        return float(loss.detach()) if read else loss

    def timed(src, read):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            step(src, read)
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / a.steps
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps, host_ms

    for _ in range(3):
        step(resident, False)
        step(pinned, True)
    step(u8, True)
    for rep in range(2):
        for name, src, read in (("resident, no read", resident, False), ("resident, float(loss)", resident, True),
                                ("pinned fp32, no read", pinned, False), ("pinned fp32, float(loss)", pinned, True),
                                ("pinned uint8, float(loss)", u8, True)):
            ms, hms = timed(src, read)
            print("%-28s %7.3f ms/step   host enqueue %7.3f ms/step" % (name, ms, hms))


if __name__ == "__main__":
    main()
