"""Per-bucket timeline of the native gradient reduction inside one training step (multi-GPU): when each bucket became ready
(its last writer was issued on the main stream), when its NVLS reduction started and finished on the sync stream, and how
long the main stream waited for the reductions after its last backward kernel (the EXPOSED synchronisation time).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sync_timeline.py [--batch 256] [--heads]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--heads", action="store_true")
    ap.add_argument("--model", default="vitb16")
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from segclip_b200.config import shape_state_dict
    from segclip_b200.modeling import SegCLIP
    from segclip_b200.p2p import EmbeddingExchange
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = bench.model_config(a)
    tc = argparse.Namespace(local_rank=local, rank=rank, world_size=world, first_stage_layer=10, use_vision_mae_recon=a.heads,
                            use_seglabel=a.heads, precision=a.precision)
    torch.manual_seed(0)
    model = SegCLIP(shape_state_dict(cfg), tc).to(dev).train()
    model.attach_exchange(EmbeddingExchange(dist.group.WORLD, dev))
    model.enable_native_grad_sync(dist.group.WORLD)
    src = {k: v.to(dev) for k, v in bench.synthetic_batch(cfg, a.batch, 0, rank, a.heads).items()}

    def step():
        model.zero_grad(set_to_none=True)
        loss = model(src["input_ids"], None, None, src["image"], image_seg=src["image_seg"] if a.heads else None)
        e_fwd = torch.cuda.Event(enable_timing=True)
        e_fwd.record()
        loss.backward()
        return e_fwd

    for _ in range(4):
        step()
    nv = model._engine.nvls
    if nv is None:
        if rank == 0:
            print("NVLS gradient sync not active on this system (NCCL buckets): no per-bucket timeline")
        dist.barrier()
        return
    rows = []
    for rep in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        nv.profile = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e_fwd = step()
        e1.record()
        torch.cuda.synchronize()
        prof, nv.profile = nv.profile, None
        rows.append((e0, e_fwd, e1, prof))
    e0, e_fwd, e1, prof = rows[-1]
    out = ["# rank %d of %d, batch %d/GPU, heads=%s: one step = %.3f ms (forward %.3f ms); times in ms from step start"
           % (rank, world, a.batch, a.heads, e0.elapsed_time(e1), e0.elapsed_time(e_fwd)),
           "# bucket   MB      ready    start      end   (reduce ms)   GB/s"]
    k = 0
    for kind, nbytes, ev in prof:
        if kind == "bucket":
            t = [e0.elapsed_time(x) for x in ev]
            out.append("  %4d  %6.1f  %8.3f %8.3f %8.3f   %8.3f   %7.1f" % (k, nbytes / 1e6, t[0], t[1], t[2], t[2] - t[1], nbytes / (t[2] - t[1]) / 1e6))
            k += 1
        else:
            t = [e0.elapsed_time(x) for x in ev]
            out.append("# last backward kernel issued/finished on the main stream at %.3f ms; reductions joined at %.3f ms -> EXPOSED %.3f ms"
                       % (t[0], t[1], t[1] - t[0]))
    exposed = [[e0_.elapsed_time(p[-1][2][1]) - e0_.elapsed_time(p[-1][2][0]) for (e0_, _, _, p) in rows]]
    out.append("# exposed time over the 3 profiled steps: " + ", ".join("%.3f" % x for x in exposed[0]) + " ms")
    for r in range(world):
        dist.barrier()
        if r == rank and rank in (0, world - 1):
            print("\n".join(out), flush=True)
    dist.barrier()


if __name__ == "__main__":
    main()
