"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference (or its untouched copy oracle/_ref).

Used (a) by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and (b) by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent, i.e. on the GPU box) to pin ``oracle.segclip_oracle``
against the real code.  Nothing here is copied from the reference; it only
installs the six import/compat shims listed in SURVEY.md section 8(c):

 1. ``diffdist`` stub: differentiable all_gather (fwd dist.all_gather, bwd per-rank
    dist.reduce SUM)                                   modules/util_module.py:24,189
 2. ``boto3`` / ``botocore`` stubs                     modules/file_utils.py:20-21
 3. ``np.float`` / ``np.long`` aliases                 modules/module_mae.py:97
 4. pre-seeded ``util.logger_initialized["seg"]``      util.py:63-67
 5. a gloo process group (``barrier()`` at modules/modeling.py:354)
 6. torch-1.8 K/V flat re-interpretation inside multi_head_attention_forward when
    key batch != query batch (modules/module_seg_vit.py:213-215; README pins 1.8.0)
"""
from __future__ import annotations

import argparse
import contextlib
import logging
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

def _find_reference():
    """The reference tree: $SEGCLIP_REFERENCE, the container's /root/reference, or the untouched copy that
    oracle/make_ref.py placed under oracle/_ref (the only one that exists on the GPU box)."""
    cands = [os.environ.get("SEGCLIP_REFERENCE"), "/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "modules", "modeling.py")):
            return c
    return "/root/reference"


REF_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "modules", "modeling.py"))


_KV_LAYOUT = {"mode": "torch18_flat"}
_installed = False


def _install_shims():
    global _installed
    if _installed:
        return
    # 3. numpy aliases
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "long"):
        np.long = np.int64

    # 2. boto3 / botocore
    if "boto3" not in sys.modules:
        sys.modules["boto3"] = types.ModuleType("boto3")
    if "botocore" not in sys.modules:
        bc = types.ModuleType("botocore")
        bce = types.ModuleType("botocore.exceptions")

        class ClientError(Exception):
            pass

        bce.ClientError = ClientError
        bc.exceptions = bce
        sys.modules["botocore"] = bc
        sys.modules["botocore.exceptions"] = bce

    # 1. diffdist
    class _AllGather(torch.autograd.Function):
        @staticmethod
        def forward(ctx, tensor, world):
            outs = [torch.zeros_like(tensor) for _ in range(world)]
            dist.all_gather(outs, tensor.contiguous())
            return tuple(outs)

        @staticmethod
        def backward(ctx, *grads):
            rank = dist.get_rank()
            mine = None
            for r, g in enumerate(grads):
                g = g.contiguous().clone()
                dist.reduce(g, r, op=dist.ReduceOp.SUM)
                if r == rank:
                    mine = g
            return mine, None

    dd = types.ModuleType("diffdist")
    ddf = types.ModuleType("diffdist.functional")

    def all_gather(gather_list, tensor, group=None, next_backprop=None, inplace=True):
        return list(_AllGather.apply(tensor, len(gather_list)))

    ddf.all_gather = all_gather
    dd.functional = ddf
    sys.modules["diffdist"] = dd
    sys.modules["diffdist.functional"] = ddf

    # 6. torch-1.8 flat K/V semantics
    orig_mha = F.multi_head_attention_forward

    def patched_mha(query, key, value, *args, **kwargs):
        if key.dim() == 3 and query.dim() == 3 and key.shape[1] != query.shape[1]:
            bsz, emb = query.shape[1], query.shape[2]
            if _KV_LAYOUT["mode"] == "torch18_flat":
                key = key.contiguous().view(-1, bsz, emb)
                value = value.contiguous().view(-1, bsz, emb)
            else:  # "per_sample": what the authors presumably intended
                key = key.permute(1, 0, 2)
                value = value.permute(1, 0, 2)
        return orig_mha(query, key, value, *args, **kwargs)

    F.multi_head_attention_forward = patched_mha
    torch.nn.functional.multi_head_attention_forward = patched_mha

    # 4. logger
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import util as ref_util  # noqa: the reference's util.py

    lg = logging.getLogger("seg_oracle")
    lg.setLevel(logging.ERROR)
    ref_util.logger_initialized["seg"] = lg
    _installed = True


def init_dist(rank=0, world=1, port=29533):
    """5. gloo group (barrier() is called even for world_size 1)."""
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)


def fake_clip_state_dict(cfg):
    """Shape-only CLIP state dict: the reference reads shapes from it."""
    vw, tw, e = cfg["vision_width"], cfg["text_width"], cfg["embed_dim"]
    p, g = cfg["patch"], cfg["grid"]
    sd = {
        "visual.conv1.weight": torch.zeros(vw, 3, p, p),
        "visual.positional_embedding": torch.zeros(g * g + 1, vw),
        "visual.proj": torch.zeros(vw, e),
        "text_projection": torch.zeros(tw, e),
        "positional_embedding": torch.zeros(cfg["context"], tw),
        "token_embedding.weight": torch.zeros(cfg["vocab"], tw),
        "ln_final.weight": torch.zeros(tw),
    }
    for i in range(12):
        sd[f"visual.transformer.resblocks.{i}.attn.in_proj_weight"] = torch.zeros(1)
    for i in range(cfg["text_layers"]):
        sd[f"transformer.resblocks.{i}.ln_1.weight"] = torch.zeros(1)
    return sd


def build_reference_model(cfg, rank=0, world=1, kv_layout="torch18_flat"):
    _install_shims()
    init_dist(rank, world)
    _KV_LAYOUT["mode"] = kv_layout
    from modules.modeling import SegCLIP  # the reference, unmodified

    args = argparse.Namespace(
        local_rank=1,  # silence show_log
        rank=rank, world_size=world, pretrained_clip_name="ViT-B/16",
        first_stage_layer=cfg["first_stage_layer"],
        use_vision_mae_recon=cfg["use_mae"], use_text_mae_recon=False,
        use_seglabel=cfg["use_kl"], max_words=cfg["context"],
        mae_vis_mask_ratio=cfg.get("mae_vis_mask_ratio", 0.75), mae_seq_mask_ratio=0.15)
    model = SegCLIP(fake_clip_state_dict(cfg), args).float().train()
    return model


@contextlib.contextmanager
def injected_rand(queue):
    """F7: replay explicit uniform noise for the reference's torch.rand draws, in order."""
    orig = torch.rand
    q = list(queue)

    def fake_rand(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        t = q.pop(0)
        assert tuple(t.shape) == shape, (tuple(t.shape), shape)
        return t.clone()

    torch.rand = fake_rand
    try:
        yield
    finally:
        torch.rand = orig
        assert not q, "unused injected noise: %d tensors left" % len(q)


def run_reference(model, batch, noise, use_mae, kv_layout="torch18_flat"):
    """loss, {name: grad} from the unmodified reference on CPU."""
    _KV_LAYOUT["mode"] = kv_layout
    model.zero_grad(set_to_none=True)
    q = [noise["u1"]]
    if use_mae:
        q += [noise["u2"], noise["u3"]]
    ids = batch["input_ids"]
    with injected_rand(q):
        loss = model(ids, torch.zeros_like(ids), batch["attention_mask"], batch["image"],
                     image_seg=batch["image_seg"])
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return loss.detach(), grads
