"""bench.py -- image-text pairs/s, forward+backward, of the SegCLIP hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--heads]

Own arm : the drop-in SegCLIP module (segclip_b200) in bf16, ViT-B/16 + text-77, per-GPU batch 256
          (BASELINE.json configs[1]); weak scaling for N > 1 (one process per GPU under torchrun).
          `value`   = device-timed (CUDA events) pairs/s with the batch already resident in HBM;
          `e2e`     = the same metric through the public module API with HOST (pinned) input buffers,
                      H2D copies of ids/image and a D2H read of the loss inside the timed region.
Reference arm (--impl reference): the CPU oracle restatement of the reference's own PyTorch path
          (oracle/segclip_oracle.py; the reference itself cannot travel to the GPU box), fp32, all host
          threads, on a bounded sample (batch 16) of the same workload.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GF_PER_PAIR = {("vitb16", False): 110.0, ("vitb16", True): 144.4,      # algorithmic GEMM FLOPs, fwd+bwd = 3x fwd
               ("vitl14", False): 252.0, ("vitl14", True): 330.6}      # ViT-L/14@224 as the reference builds it (10+2 layers)
METRIC = "image-text pairs/sec (224^2, seq77) fwd+bwd"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], hbm=p["hbm_gbs"], src="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def wait_ready(self, timeout=10.0):
        """nvidia-smi takes a moment to deliver its first sample: do not start the timed region before it."""
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout:
            time.sleep(0.05)

    def mark(self, begin):
        if begin:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        rows = [r[1:] for r in self.rows if len(r) >= 9 and (self.t0 is None or r[0] >= self.t0) and
                (self.t1 is None or r[0] <= self.t1 + 0.15)]
        sm = sorted(int(float(r[1])) for r in rows)
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=int(float(rows[0][2])), reasons=reasons, samples=len(sm))


def synthetic_batch(cfg, B, seed, rank, heads):
    """SURVEY 8(d) generator (same as the oracle's make_batch, but vectorised for large B)."""
    g = torch.Generator().manual_seed(seed + rank)
    T, V, L = cfg["context"], cfg["vocab"], cfg["grid"] ** 2
    res = cfg["patch"] * cfg["grid"]
    image = torch.randn(B, 1, 3, res, res, generator=g)
    n = torch.randint(5, T - 1, (B,), generator=g)
    ids = torch.randint(1, V - 2, (B, T), generator=g)
    pos = torch.arange(T).unsqueeze(0)
    ids = torch.where(pos <= n.unsqueeze(1), ids, torch.zeros_like(ids))
    ids[:, 0] = V - 2
    ids[torch.arange(B), n + 1] = V - 1
    ids = ids.view(B, 1, T)
    seg = torch.randint(0, 6, (B, 1, cfg["grid"], cfg["grid"]), generator=g)
    return dict(input_ids=ids, attention_mask=(ids != 0).long(), image=image, image_seg=seg)


def _cpu_arm(args, steps, warmup):
    """One CPU measurement of the hot path on all host cores, fp32, batch args.cpu_batch of the same synthetic workload:
    the UNMODIFIED reference's own nn.Module (oracle/_ref, kind "reference") when it is available, else the oracle port
    (kind "port").  Returns (ms per step list, kind)."""
    from oracle import ref_harness as rh
    from oracle import segclip_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = model_config(args)
    Bc = args.cpu_batch
    batch, noise = so.make_batch(cfg, Bc, seed=0)
    kind = "port"
    model = None
    if rh.reference_available() and not os.environ.get("SEGCLIP_CPU_ARM_PORT"):
        try:
            os.environ["MASTER_ADDR"] = "127.0.0.1"
            os.environ["MASTER_PORT"] = str(29800 + os.getpid() % 150)      # not torchrun's rendezvous port
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                raise RuntimeError("process group already initialised")
            torch.manual_seed(0)
            model = rh.build_reference_model(cfg)
            kind = "reference"
        except Exception as e:          # fall back to the port, say so
            sys.stderr.write("reference arm unavailable (%s: %s); timing the oracle port\n" % (type(e).__name__, e))
            model = None
    params = so.init_params(cfg, seed=0) if model is None else None
    frozen = ("vis_mae_decoder.decoder_pos_embed",)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if model is not None:
            rh.run_reference(model, batch, noise, cfg["use_mae"])
        else:
            so.loss_and_grads(params, batch, noise, cfg, "torch18_flat", frozen=frozen)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, kind, cores


def run_reference(args):
    """CPU arm: the reference's own PyTorch path (or the oracle port when the reference copy is absent), fp32, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Bc = args.cpu_batch
    times, kind, cores = _cpu_arm(args, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    v = Bc / (ms / 1e3)
    what = "the unmodified reference nn.Module (oracle/_ref + 6 import shims)" if kind == "reference" else "oracle port"
    sample = "%s, batch %d of the same synthetic workload, fp32, %d steps after %d warm-up" % (what, Bc, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "cpu_batch": Bc},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def model_config(args):
    from segclip_b200 import config
    if args.model == "vitl14":      # BASELINE configs[4]: width 1024 / 16 heads / patch 14, text 768 / 12 heads, E = 768 (F5: 10+2 layers)
        return config.vit_l14(use_mae=args.heads, use_kl=args.heads)
    return config.vit_b16(use_mae=args.heads, use_kl=args.heads)


def workload_name(args):
    name = "ViT-B/16 SegCLIP" if args.model == "vitb16" else "ViT-L/14 (10+2 layers, as the reference builds it) SegCLIP"
    tag = "BASELINE configs[1]" if (args.model == "vitb16" and not args.heads) else \
        ("BASELINE configs[3] shape" if args.model == "vitb16" else "BASELINE configs[4] shape")
    return "%s full fwd+bwd, per-GPU batch %d, %s, bf16 (%s)" % (
        name, args.batch, "contrastive + MAE-recon + superpixel-KL" if args.heads else "contrastive only", tag)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--heads", action="store_true", help="enable MAE-reconstruction + superpixel-KL heads (config 4)")
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--model", default="vitb16", choices=["vitb16", "vitl14"])
    ap.add_argument("--ddp", action="store_true", help="use torch DDP for the gradient mean instead of the native overlapped sync")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from segclip_b200 import _lib
    from segclip_b200.config import shape_state_dict
    from segclip_b200.modeling import SegCLIP

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = model_config(args)
    tc = argparse.Namespace(local_rank=local, rank=rank, world_size=world, first_stage_layer=10,
                            use_vision_mae_recon=args.heads, use_seglabel=args.heads, precision=args.precision)
    torch.manual_seed(0)
    model = SegCLIP(shape_state_dict(cfg), tc).to(dev).train()
    net = model
    if world > 1:
        from segclip_b200.p2p import EmbeddingExchange
        model.attach_exchange(EmbeddingExchange(dist.group.WORLD, dev))
        if args.ddp:          # comparator: stock DistributedDataParallel hooks (all-reduce after the native backward)
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True,
                                                            gradient_as_bucket_view=True)
        else:                 # product path: bucketed all-reduce overlapped inside the native backward tape
            model.enable_native_grad_sync(dist.group.WORLD)
    B = args.batch
    host = synthetic_batch(cfg, B, 0, rank, args.heads)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}

    def step(src, read_loss):
        net.zero_grad(set_to_none=True)
        loss = net(src["input_ids"], None, None, src["image"], image_seg=src["image_seg"] if args.heads else None)
        loss.backward()
        return float(loss.detach()) if read_loss else loss

    def fwd_only(src, read_loss):
        return net(src["input_ids"], None, None, src["image"], image_seg=src["image_seg"] if args.heads else None)

    def timed(src, read_loss, steps, fn=None):
        fn = fn or step
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(src, read_loss)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    loss0 = None
    for i in range(args.warmup):
        l_ = step(resident, False)
        if i == 0:
            loss0 = float(l_.detach())          # step-0 loss of this rank (random-init weights, synthetic batch)
    if rank == 0:
        sampler.wait_ready()
        sampler.mark(True)
    l0 = _lib.launch_count()
    ms = timed(resident, False, args.steps)
    launches = (_lib.launch_count() - l0) // args.steps
    step(pinned, True)
    ms_e2e = timed(pinned, True, args.steps)          # clocks are sampled over both timed regions
    # the same end-to-end step fed raw uint8 pixels (the library's optional input boundary, SURVEY 8(f) rank 3: 1 byte per pixel
    # over PCIe, CLIP mean / std normalisation on the GPU) -- informational, the headline `e2e` is the reference's fp32 contract
    pinned_u8 = dict(pinned)
    pinned_u8["image"] = (torch.rand(host["image"].shape) * 255).to(torch.uint8).pin_memory()
    step(pinned_u8, True)
    ms_e2e_u8 = timed(pinned_u8, True, args.steps)
    ms_fwd = timed(resident, False, args.steps, fwd_only)      # forward tape only (north_star: >= 70 % on the forward)
    if world > 1:
        lt = torch.tensor([loss0], device=dev)
        lall = [torch.zeros_like(lt) for _ in range(world)]
        dist.all_gather(lall, lt)
        losses = [float(x) for x in lall]
    else:
        losses = [loss0]
    if rank == 0:
        sampler.mark(False)
    clocks = sampler.stop() if rank == 0 else None

    # dominant kernel (tcgen05 GEMM): achieved TFLOP/s over all its launches of one step, CUDA events
    gemm = model._engine.profile_gemm(B) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    pk = peaks()
    value = B * world / (ms / 1e3)
    gf = GF_PER_PAIR[(args.model, args.heads)]
    out = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "per_gpu_batch": B, "global_batch": B * world,
                   "parallelism": "dp%d" % world, "grad_sync": ("ddp" if args.ddp else ("native-overlapped, NVLS multimem kernel" if model._engine.nvls is not None else "native-overlapped, NCCL buckets")) if world > 1 else "none", "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2"},
        "e2e": {"value": B * world / (ms_e2e / 1e3), "unit": "pairs/s",
                "h2d_bytes_per_step": int(host["input_ids"].numel() * 8 + host["image"].numel() * 4 +
                                          (host["image_seg"].numel() * 8 if args.heads else 0)),
                "d2h_bytes_per_step": 4},
        "e2e_uint8_input": {"value": B * world / (ms_e2e_u8 / 1e3), "unit": "pairs/s",
                            "h2d_bytes_per_step": int(host["input_ids"].numel() * 8 + host["image"].numel() +
                                                      (host["image_seg"].numel() * 8 if args.heads else 0)),
                            "note": "optional uint8 image boundary (normalised on the device); not the headline"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "loss_step0": losses, "loss_finite": all(x == x and abs(x) != float("inf") for x in losses),
        "fwd_tensor_frac": {"ms_fwd": ms_fwd, "achieved_tflops_per_gpu": B * gf / 3.0 / ms_fwd, "peak": pk["tflops"],
                            "frac": B * gf / 3.0 / ms_fwd / pk["tflops"], "flops_per_pair_gf": gf / 3.0},
        "step_tensor_frac": {"achieved_tflops_per_gpu": value / world * gf / 1e3, "peak": pk["tflops"],
                             "frac": value / world * gf / 1e3 / pk["tflops"], "flops_per_pair_gf": gf},
        "roofline": {"bound": "tensor", "achieved": gemm["tflops"], "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": gemm["tflops"] / pk["tflops"], "traffic": gemm_traffic(args),
                     "kernel": "gemm_tc2_kernel (tcgen05 cta_group::2, all launches of one step)",
                     "launches_per_step": gemm["launches"], "share_of_step": gemm["ms"] / ms, "peak_source": pk["src"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(out))
    if world > 1:
        dist.barrier()


def gemm_traffic(args):
    """Average DRAM bytes (read + write) per gemm_tc2_kernel launch of this workload, from the committed ncu pass
    (profiles/r1_gemm_traffic.json, written by tools/summarize_launches.py --traffic); None for other workloads."""
    path = os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
    if args.model != "vitb16" or args.heads or args.batch != 256 or not os.path.exists(path):
        return None
    return json.load(open(path)).get("dram_bytes_per_launch")


def cpu_baseline(args):
    """The only leg of the own arm that touches oracle/: the CPU path timed beside the GPU number (bounded sample)."""
    times, kind, cores = _cpu_arm(args, 2, 1)
    best = min(times)
    what = "unmodified reference nn.Module (oracle/_ref)" if kind == "reference" else "oracle port"
    return {"value": args.cpu_batch / best, "unit": "pairs/s", "cores": cores, "kind": kind,
            "sample": "%s (fp32 PyTorch CPU), batch %d, best of 2 after 1 warm-up" % (what, args.cpu_batch)}


if __name__ == "__main__":
    main()
