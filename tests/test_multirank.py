"""Multi-rank path.

CPU (gloo, world_size 2): the library-collective comparator exchange fills the gathered buffers in rank order.
GPU (needs >= 2 devices, `gpurun --gpus 2`): two ranks of the real CUDA path -- once with the NVLink P2P write
kernel and once with the NCCL comparator -- reproduce the 2-rank fixture of the unmodified reference (diffdist
all-gather + DDP gradient mean, tests/golden/toy_heads_flat_w2.json); `gpurun --gpus 8`: the 8-rank fixture."""
import argparse
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden_util import compare_grads, load_case


def _cpu_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from segclip_b200.p2p import EmbeddingExchange
    ex = EmbeddingExchange(dist.group.WORLD, "cpu")
    B, E = 3, 8
    t_all, v_all, lse_all = ex.buffers(B, E)
    lo = rank * B
    t_all[lo:lo + B] = rank + 1
    v_all[lo:lo + B] = 10 * (rank + 1)
    lse_all[0, lo:lo + B] = 100 + rank
    lse_all[1, lo:lo + B] = 200 + rank
    ex.gather_embeddings()
    ex.gather_lse()
    ex.release()
    ok = all(float(t_all[r * B:(r + 1) * B].mean()) == r + 1 and float(v_all[r * B:(r + 1) * B].mean()) == 10 * (r + 1) and
             float(lse_all[0, r * B]) == 100 + r and float(lse_all[1, r * B + B - 1]) == 200 + r for r in range(world))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_collective_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, 29561, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get() for _ in range(2))
    for p in procs:
        p.join()
    assert res == {0: True, 1: True}


def _gpu_worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["SEGCLIP_EXCHANGE"] = "nccl" if mode == "nccl" else "p2p"
    os.environ["SEGCLIP_P2P_TIMEOUT_S"] = "60"       # a protocol bug must end the test, not hang the box
    if mode == "native_nccl":
        os.environ["SEGCLIP_GRAD_SYNC"] = "nccl"     # comparator transport of the gradient buckets
    elif mode == "native_nvls":
        os.environ["SEGCLIP_GRAD_SYNC"] = "nvls"     # the library's multimem kernel; fails loudly without NVSwitch multicast
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import ref_harness as rh
    from oracle import segclip_oracle as so
    from segclip_b200.engine import FROZEN_STEM
    from segclip_b200.modeling import SegCLIP
    from segclip_b200.p2p import EmbeddingExchange
    g = load_case("toy_heads_flat_w%d" % world)
    cfg = g["config"]
    args = argparse.Namespace(local_rank=rank, rank=rank, world_size=world, first_stage_layer=cfg["first_stage_layer"],
                              use_vision_mae_recon=True, use_seglabel=True, precision="fp32", kv_layout=g["kv_layout"])
    model = SegCLIP(rh.fake_clip_state_dict(cfg), args)
    model.load_state_dict(so.init_params(cfg, seed=g["param_seed"]), strict=False)
    model = model.to(dev).train()
    model.attach_exchange(EmbeddingExchange(dist.group.WORLD, dev))
    if mode.startswith("native"):           # gradient mean inside the native backward (overlapped bucket reductions)
        model.enable_native_grad_sync(dist.group.WORLD)
    batch, noise = so.make_batch(cfg, g["batch"], seed=g["batch_seed"], rank=rank)
    model.inject_noise({k: v.to(dev) for k, v in noise.items()})
    ids = batch["input_ids"]
    losses = []
    for step in range(2):                   # two steps: exercises the epoch / release protocol
        model.zero_grad(set_to_none=True)
        loss = model(ids, torch.zeros_like(ids), batch["attention_mask"], batch["image"], image_seg=batch["image_seg"])
        loss.backward()
        losses.append(float(loss.detach()))
        if mode == "p2p_eval" and step == 0 and rank == 0:
            # rank-0-only evaluation between two training steps, with ANOTHER batch size (main_task_align.py:484-490):
            # must neither enter a collective nor re-point the exchange buffers of the training plan, while the other
            # rank is already waiting inside the next step's exchange
            model.eval()
            with torch.no_grad():
                model.clip.encode_text(ids[:1, 0])
                model.clip.encode_image(batch["image"][:1, 0])
            model.train()
    grads = {}
    for n, p in model.named_parameters():
        if p.grad is not None:
            gr = p.grad.detach().clone()
            if not mode.startswith("native"):
                dist.all_reduce(gr)
                gr = gr / world
            grads[n] = gr.cpu()
    bad = compare_grads(grads, g["grads"], tol=1e-3, skip=FROZEN_STEM) if rank == 0 else []
    if mode == "native_nvls":
        assert model._engine.nvls is not None
        model._engine.nvls.check()
    q.put((rank, losses, g["loss"][rank], bad[:5]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world,mode", [(2, "p2p"), (2, "nccl"), (2, "native"), (2, "native_nccl"), (2, "native_nvls"),
                                        (2, "p2p_eval"), (8, "native"), (8, "native_nvls")])
def test_ranks_match_reference_fixture(world, mode):
    """W ranks of the real CUDA path against the W-rank fixture of the unmodified reference (W gloo processes:
    diffdist all-gather + DDP gradient mean).  W = 8 checks the epoch / consumed flag protocol at the full fan-out of the box."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = {"p2p": 29571, "nccl": 29573, "native": 29575, "p2p_eval": 29577, "native_nccl": 29579, "native_nvls": 29581}[mode] + 20 * (world == 8)
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get() for _ in range(world)]
    for p in procs:
        p.join(120)
    for rank, losses, want, bad in res:
        for l in losses:
            assert abs(l - want) <= 1e-3 * abs(want), (rank, losses, want)
        assert not bad, bad


def _nvls_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["SEGCLIP_GRAD_SYNC"] = "nvls"
    os.environ["SEGCLIP_P2P_TIMEOUT_S"] = "60"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from segclip_b200.allreduce import NvlsGradSync
    nv = NvlsGradSync.create(dist.group.WORLD, dev)
    n = 3 * 1024 * 1024 + 8
    buf = nv.alloc(n)
    assert buf is not None, "no NVSwitch multicast mapping"
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ok = True
    cur = torch.cuda.current_stream(dev)
    for (lo, hi) in ((0, n), (4, 1028), (1024 * 1024 + 4, n - 4), (8, 12)):      # whole buffer, tiny, unaligned-to-slices ranges
        buf.copy_(torch.randn(n, device=dev, generator=g))
        want = buf.clone()
        dist.all_reduce(want[lo:hi])
        want[lo:hi] /= world
        torch.cuda.synchronize()
        dist.barrier()
        nv.all_reduce(lo, hi, cur)
        nv.join(cur)
        torch.cuda.synchronize()
        nv.check()
        err = float((buf - want).abs().max())
        ok = ok and err <= 1e-5 * (1 + float(want.abs().max()))
        dist.barrier()
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 8])
def test_nvls_allreduce_matches_nccl(world):
    """sc_nvls_allreduce (multimem.ld_reduce / multimem.st two-shot kernel on the symmetric gradient buffer) against NCCL's
    all-reduce on sub-ranges of the buffer; values outside the range must stay untouched."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_nvls_worker, args=(r, world, 29641 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(120)
    assert all(res.values()), res
