"""Drop-in ``SegCLIP`` module: the reference's nn.Module surface over the B200-native engine.

Mirrors modules/modeling.py (reference): ``SegCLIP.from_pretrained(state_dict=, task_config=)`` (:27),
``SegCLIP(clip_state_dict, task_config)`` (:79) and ``forward(input_ids, token_type_ids,
attention_mask, image, image_seg=None)`` (:174) -> scalar loss in training mode, ``None`` in eval.
The parameter tree (names, shapes, nn.Parameter identity per name) is the
reference's, so ``prep_optimizer`` / the freeze loop / checkpoints of main_task_align.py work
unchanged.  All arithmetic runs in libsegclip_b200.so; this file only owns parameters, marshals
inputs and hooks the native forward/backward into autograd.
"""
import math
import os

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import ops
from .engine import DEC_DEPTH, FROZEN_STEM, TRAINABLE_STEM, Engine, G


def get_attr(cfg, name, default):
    """modules/util_module.py:206-211"""
    return getattr(cfg, name) if hasattr(cfg, name) else default


# ----------------------------------------------------------------------------------------------
# parameter tree
# ----------------------------------------------------------------------------------------------
class _Node(nn.Module):
    """Bare container; children and parameters are attached by dotted name."""


def _attach(root, dotted, param):
    mod = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], param)


def _sincos_2d(dim, grid):
    """Fixed decoder position table (modules/module_mae.py:63-108): first half of the channels
    encodes the W coordinate, second half the H coordinate, each as [sin | cos]."""
    def one(d, pos):
        omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=np.float64) / (d / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    gw, gh = np.meshgrid(np.arange(grid, dtype=np.float32), np.arange(grid, dtype=np.float32))
    emb = np.concatenate([one(dim // 2, gw), one(dim // 2, gh)], axis=1)
    return np.concatenate([np.zeros([1, dim]), emb], axis=0)


def _param_specs(cfg):
    """(name, shape, init) for every parameter, in the reference's registration spirit.
    init: ("normal", std) | ("uniform", bound) | ("const", value) | ("xavier",) | ("sincos",)"""
    vw, tw, e = cfg["vision_width"], cfg["text_width"], cfg["embed_dim"]
    ps, gr, fsl = cfg["patch"], cfg["grid"], cfg["first_stage_layer"]
    specs = []

    def add(name, shape, init):
        specs.append((name, tuple(shape), init))

    def clip_block(pre, d, in_proj_init):
        add(pre + "attn.in_proj_weight", (3 * d, d), in_proj_init)
        add(pre + "attn.in_proj_bias", (3 * d,), ("const", 0.0))
        add(pre + "attn.out_proj.weight", (d, d), ("normal", 0.02))      # util_module.py:70-85
        add(pre + "attn.out_proj.bias", (d,), ("const", 0.0))
        add(pre + "ln_1.weight", (d,), ("const", 1.0))
        add(pre + "ln_1.bias", (d,), ("const", 0.0))
        add(pre + "mlp.c_fc.weight", (4 * d, d), ("normal", 0.02))
        add(pre + "mlp.c_fc.bias", (4 * d,), ("const", 0.0))
        add(pre + "mlp.c_proj.weight", (d, 4 * d), ("normal", 0.02))
        add(pre + "mlp.c_proj.bias", (d,), ("const", 0.0))
        add(pre + "ln_2.weight", (d,), ("const", 1.0))
        add(pre + "ln_2.bias", (d,), ("const", 0.0))

    v = "clip.visual."
    add(v + "class_embedding", (vw,), ("normal", vw ** -0.5))
    add(v + "positional_embedding", (gr * gr + 1, vw), ("normal", vw ** -0.5))
    add(v + "proj", (vw, e), ("normal", vw ** -0.5))
    add(v + "conv1.weight", (vw, 3, ps, ps), ("uniform", (3 * ps * ps) ** -0.5))
    add(v + "ln_pre.weight", (vw,), ("const", 1.0))
    add(v + "ln_pre.bias", (vw,), ("const", 0.0))
    t = v + "transformer."
    for i in range(fsl):
        clip_block(f"{t}layers0.{i}.", vw, ("xavier",))
    s = t + "semantic_layer2."
    add(s + "semantic_center", (G, vw), ("normal", 0.02))
    add(s + "norm.weight", (vw,), ("const", 1.0))
    add(s + "norm.bias", (vw,), ("const", 0.0))
    for i in range(2):
        c = f"{s}cross_att.{i}."
        add(c + "attn.in_proj_weight", (3 * vw, vw), ("xavier",))
        add(c + "attn.in_proj_bias", (3 * vw,), ("const", 0.0))
        add(c + "attn.out_proj.weight", (vw, vw), ("normal", 0.02))
        add(c + "attn.out_proj.bias", (vw,), ("const", 0.0))
        for ln in ("ln_x", "ln_k"):
            add(c + ln + ".weight", (vw,), ("const", 1.0))
            add(c + ln + ".bias", (vw,), ("const", 0.0))
        add(c + "mlp.c_fc.weight", (4 * vw, vw), ("normal", 0.02))
        add(c + "mlp.c_fc.bias", (4 * vw,), ("const", 0.0))
        add(c + "mlp.c_proj.weight", (vw, 4 * vw), ("normal", 0.02))
        add(c + "mlp.c_proj.bias", (vw,), ("const", 0.0))
        add(c + "ln_2.weight", (vw,), ("const", 1.0))
        add(c + "ln_2.bias", (vw,), ("const", 0.0))
    add(s + "cross_ln.weight", (vw,), ("const", 1.0))
    add(s + "cross_ln.bias", (vw,), ("const", 0.0))
    add(s + "k_conv.weight", (vw, 64, 1), ("uniform", 64 ** -0.5))
    add(s + "k_ln.weight", (vw,), ("const", 1.0))
    add(s + "k_ln.bias", (vw,), ("const", 0.0))
    add(s + "v_conv.weight", (vw, 64, 1), ("uniform", 64 ** -0.5))
    add(s + "proj_o.ln.weight", (vw,), ("const", 1.0))
    add(s + "proj_o.ln.bias", (vw,), ("const", 0.0))
    add(s + "proj_o.mlp.fc1.weight", (4 * vw, vw), ("normal", 0.02))
    add(s + "proj_o.mlp.fc1.bias", (4 * vw,), ("const", 0.0))
    add(s + "proj_o.mlp.fc2.weight", (vw, 4 * vw), ("normal", 0.02))
    add(s + "proj_o.mlp.fc2.bias", (vw,), ("const", 0.0))
    for i in range(12 - fsl):
        clip_block(f"{t}layers2.{i}.", vw, ("xavier",))
    for i in range(12 - fsl):
        clip_block(f"{t}layers_mae2.{i}.", vw, ("xavier",))
    r = t + "reconstruct_layer2.rec_proj_a.a_fc."
    add(r + "weight", (G, G), ("normal", 0.02))
    add(r + "bias", (G,), ("const", 0.0))
    add(v + "ln_post.weight", (vw,), ("const", 1.0))
    add(v + "ln_post.bias", (vw,), ("const", 0.0))
    for i in range(cfg["text_layers"]):
        clip_block(f"clip.transformer.resblocks.{i}.", tw, ("normal", tw ** -0.5))
    add("clip.token_embedding.weight", (cfg["vocab"], tw), ("normal", 0.02))
    add("clip.positional_embedding", (cfg["context"], tw), ("normal", 0.01))
    add("clip.ln_final.weight", (tw,), ("const", 1.0))
    add("clip.ln_final.bias", (tw,), ("const", 0.0))
    add("clip.text_projection", (tw, e), ("normal", tw ** -0.5))
    add("clip.logit_scale", (), ("const", math.log(1 / 0.07)))
    if cfg["use_mae"]:
        dd = vw // 2
        m = "vis_mae_decoder."
        add(m + "mask_token", (1, 1, dd), ("normal", 0.02))
        add(m + "decoder_pos_embed", (1, gr * gr + 1, dd), ("sincos",))
        add(m + "decoder_embed.weight", (dd, vw), ("normal", 0.02))
        add(m + "decoder_embed.bias", (dd,), ("const", 0.0))
        for i in range(DEC_DEPTH):
            b = f"{m}decoder_blocks.{i}."
            add(b + "norm1.weight", (dd,), ("const", 1.0))
            add(b + "norm1.bias", (dd,), ("const", 0.0))
            add(b + "attn.qkv.weight", (3 * dd, dd), ("normal", 0.02))
            add(b + "attn.qkv.bias", (3 * dd,), ("const", 0.0))
            add(b + "attn.proj.weight", (dd, dd), ("normal", 0.02))
            add(b + "attn.proj.bias", (dd,), ("const", 0.0))
            add(b + "norm2.weight", (dd,), ("const", 1.0))
            add(b + "norm2.bias", (dd,), ("const", 0.0))
            add(b + "mlp.fc1.weight", (4 * dd, dd), ("normal", 0.02))
            add(b + "mlp.fc1.bias", (4 * dd,), ("const", 0.0))
            add(b + "mlp.fc2.weight", (dd, 4 * dd), ("normal", 0.02))
            add(b + "mlp.fc2.bias", (dd,), ("const", 0.0))
        add(m + "decoder_norm.weight", (dd,), ("const", 1.0))
        add(m + "decoder_norm.bias", (dd,), ("const", 0.0))
        add(m + "decoder_pred.weight", (3 * ps * ps, dd), ("normal", 0.02))
        add(m + "decoder_pred.bias", (3 * ps * ps,), ("const", 0.0))
    return specs


def _init_tensor(shape, init, cfg):
    kind = init[0]
    if kind == "normal":
        return torch.randn(shape) * init[1]
    if kind == "uniform":
        return (torch.rand(shape) * 2 - 1) * init[1]
    if kind == "const":
        return torch.full(shape, float(init[1]))
    if kind == "xavier":
        t = torch.empty(shape)
        nn.init.xavier_uniform_(t)
        return t
    if kind == "sincos":
        return torch.from_numpy(_sincos_2d(shape[-1], cfg["grid"])).float().view(shape)
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------
# autograd bridge
# ----------------------------------------------------------------------------------------------
class _ClipFacade(_Node):
    """`model.clip`: owns the CLIP parameters (reference attribute tree) and exposes the inference entry points of
    modules/module_clip.py:89-143.  Training-mode encoding is part of the fused SegCLIP.forward()."""

    def _own(self):
        return self.__dict__["_owner"]()

    def encode_image(self, image, return_hidden=False, video_frame=-1, mask_ratio=0.):
        owner = self._own()
        if owner.training or mask_ratio > 0.:
            raise NotImplementedError("training-mode / masked encode_image is fused into SegCLIP.forward(); call model(...)")
        eng = owner._get_engine()
        image = torch.as_tensor(image)
        B = image.shape[0]
        gh, gw = image.shape[-2] // eng.patch, image.shape[-1] // eng.patch
        if image.shape[-2] != gh * eng.patch or image.shape[-1] != gw * eng.patch:
            raise ValueError("image size %s is not a multiple of the patch size %d" % (tuple(image.shape[-2:]), eng.patch))
        if (gh, gw) != (eng.grid, eng.grid):
            # another input size (zero-shot segmentation at 2x the training resolution, vit_seg.py:157): inference clone of the
            # engine with a bicubically resized positional table (module_clip_vtransformer.py:35-53)
            eng = eng.for_grid(gh, gw)
        b = eng.infer(B, image=image.to(eng.dev), norm=owner.image_norm)
        hidden = b["v.hidden9"].view(B, G + 1, eng.E).clone()
        x = hidden[:, 0]
        if not return_hidden:
            return x
        patches = b["v%d.x_out" % (eng.fsl - 1)] if eng.fsl > 0 else b["v.stem.x0"]
        idx = b["v.sem.idx"].long()
        hard = torch.zeros(B, G, eng.Lp, device=eng.dev).scatter_(1, idx.unsqueeze(1), 1.0)
        mid = {"hidden": patches.view(B, eng.Lp, eng.vw).clone(),
               "attns": [{"soft_attn": b["v.sem.soft"].clone(), "hard_attn": hard}]}
        return x, hidden, mid

    def encode_text(self, text, attn_mask=None, return_hidden=False, mask_ratio=0.):
        owner = self._own()
        if owner.training or mask_ratio > 0.:
            raise NotImplementedError("training-mode / masked encode_text is fused into SegCLIP.forward(); call model(...)")
        eng = owner._get_engine()
        text = torch.as_tensor(text)
        B = text.shape[0]
        b = eng.infer(B, ids=text.to(eng.dev))
        x = b["t.raw"].clone()
        if return_hidden:
            return x, b["t.hidden_all"].view(B, eng.Tctx, eng.E).clone()
        return x


class _NativeStep(torch.autograd.Function):
    """loss = native_forward(params...); backward replays the native backward tape."""

    @staticmethod
    def forward(ctx, owner, B, inputs, noise, forced, *params):
        eng = owner._engine
        loss = eng.forward(B, inputs, noise, forced)
        ctx.owner, ctx.B = owner, B
        return loss.clone().view(())

    @staticmethod
    def backward(ctx, gout):
        owner = ctx.owner
        eng = owner._engine
        gflat = eng.backward(ctx.B)
        out = torch.empty_like(gflat)
        scale = gout.detach().to(torch.float32).contiguous().view(1)
        if eng.sync_world > 1 and eng.nvls is None:
            scale = scale / eng.sync_world          # mean over ranks, folded into the hand-over copy (the NVLS kernel scales itself)
        ops.convert_op(gflat, out, scale)()          # grad * grad_output, read on the device
        grads = []
        for name, p in owner._active_items:
            if p.requires_grad:
                o = eng.goffs[name]
                grads.append(out[o:o + p.numel()].view(p.shape))
            else:
                grads.append(None)
        return (None, None, None, None, None) + tuple(grads)


class SegCLIP(nn.Module):
    def __init__(self, clip_state_dict, task_config):
        super().__init__()
        self.task_config = task_config
        self.ignore_image_index = -1
        sd = clip_state_dict
        assert "visual.proj" in sd, "only the ViT variants are supported (as in the reference, modeling.py:86-87)"
        vision_width = sd["visual.conv1.weight"].shape[0]
        patch = sd["visual.conv1.weight"].shape[-1]
        grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
        text_width = sd["ln_final.weight"].shape[0]
        self.cfg = dict(
            vision_width=vision_width, text_width=text_width, embed_dim=sd["text_projection"].shape[1], patch=patch,
            grid=grid, context=sd["positional_embedding"].shape[0], vocab=sd["token_embedding.weight"].shape[0],
            text_layers=len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks"))),
            first_stage_layer=get_attr(task_config, "first_stage_layer", 10),
            use_mae=bool(get_attr(task_config, "use_vision_mae_recon", False)),
            use_kl=bool(get_attr(task_config, "use_seglabel", False)))
        if get_attr(task_config, "use_text_mae_recon", False):
            raise NotImplementedError("use_text_mae_recon is outside the B200 hot path (off in the reference recipe)")
        self.first_stage_layer = self.cfg["first_stage_layer"]
        self.use_vision_mae_recon = self.cfg["use_mae"]
        self.use_seglabel = self.cfg["use_kl"]
        self.use_text_mae_recon = False
        self.vis_mask_ratio = get_attr(task_config, "mae_vis_mask_ratio", 0.75)
        if not 0.0 < self.vis_mask_ratio < 1.0:
            raise ValueError("mae_vis_mask_ratio must be in (0, 1), got %r" % (self.vis_mask_ratio,))
        if self.vis_mask_ratio != 0.75:
            self.cfg["mae_vis_mask_ratio"] = float(self.vis_mask_ratio)
        import weakref
        self.add_module("clip", _ClipFacade())
        self.clip.__dict__["_owner"] = weakref.ref(self)
        for name, shape, init in _param_specs(self.cfg):
            p = nn.Parameter(_init_tensor(shape, init, self.cfg))
            if name in FROZEN_STEM:
                p.requires_grad = False       # frozen by the reference recipe (main_task_align.py:389-441)
            _attach(self, name, p)
        self.precision = get_attr(task_config, "precision", "bf16")
        self.kv_layout = get_attr(task_config, "kv_layout", "torch18_flat")
        self._engine = None
        self._noise = None
        self._forced = None
        self._exchange = None
        self._sync_group = None
        # CLIP preprocessing constants (dataloaders/rawimage_util.py Normalize); used only for uint8 image batches
        self.image_norm = (get_attr(task_config, "image_mean", (0.48145466, 0.4578275, 0.40821073)),
                           get_attr(task_config, "image_std", (0.26862954, 0.26130258, 0.27577711)))

    # ---- reference-compatible constructors ---------------------------------------------------
    @classmethod
    def from_pretrained(cls, state_dict=None, cache_dir=None, type_vocab_size=2, *inputs, clip_state_dict=None, **kwargs):
        """modules/modeling.py:27-75.  ``clip_state_dict`` (extension) supplies the CLIP weights directly;
        otherwise the OpenAI archive is looked up on disk (no network on the training box)."""
        task_config = kwargs.get("task_config")
        if task_config is not None:
            if not hasattr(task_config, "local_rank"):
                task_config.__dict__["local_rank"] = 0
            elif task_config.local_rank == -1:
                task_config.local_rank = 0
        state_dict = {} if state_dict is None else state_dict
        if clip_state_dict is None:
            clip_state_dict = _load_clip_archive(get_attr(task_config, "pretrained_clip_name", "ViT-B/16"))
        clip_state_dict = {k: v for k, v in clip_state_dict.items()
                           if k not in ("input_resolution", "context_length", "vocab_size")}
        fsl = get_attr(task_config, "first_stage_layer", 10)
        for key, val in clip_state_dict.items():          # CLIP -> layers0/layers2 remap (:50-68)
            new_key = "clip." + key
            if "visual.transformer." in key:
                n = int(new_key.split(".")[4])
                if n >= fsl:
                    parts = new_key.replace(".resblocks.", ".layers2.").split(".")
                    parts[4] = str(n - fsl)
                    new_key = ".".join(parts)
                else:
                    new_key = new_key.replace(".resblocks.", ".layers0.")
            if new_key not in state_dict:
                state_dict[new_key] = val.clone()
        model = cls(clip_state_dict, *inputs, **kwargs)
        own = model.state_dict()
        filtered = {k: v.float() for k, v in state_dict.items() if k in own and own[k].shape == v.shape}
        model.load_state_dict(filtered, strict=False)
        return model

    # ---- engine plumbing -------------------------------------------------------------------------
    def _get_engine(self):
        named = None
        if self._engine is not None:
            # the stem the reference recipe freezes (main_task_align.py:389-441) was (un)frozen since the engine was built:
            # rebuild it with / without the stem's backward
            named = dict(self.named_parameters())
            if self._engine.train_stem != any(named[n].requires_grad for n in TRAINABLE_STEM):
                assert self._sync_group is None or not self._engine.plans, "(un)freeze the stem before the first step when the native gradient sync is on"
                self._engine = None
        if self._engine is None or self._engine.params_moved():
            named = named or dict(self.named_parameters())
            dev = next(iter(named.values())).device
            if dev.type != "cuda":
                raise L.SegclipB200Error("segclip_b200 has no CPU path: move the module to a CUDA device first")
            tc = self.task_config
            self._engine = Engine(self.cfg, named, self.precision, self.kv_layout, int(get_attr(tc, "rank", 0)),
                                  int(get_attr(tc, "world_size", 1)),
                                  train_stem=any(named[n].requires_grad for n in TRAINABLE_STEM))
            self._param_items = list(named.items())
            self._untouched = set()
            if not self.cfg["use_mae"]:
                self._untouched = {n for n in named if ".layers_mae2." in n or ".reconstruct_layer2." in n}
            # parameters that take part in this configuration's graph; the others (MAE-only branches when the MAE
            # head is off) stay outside autograd exactly like the reference's unused parameters (DDP is built with
            # find_unused_parameters=True there, main_task_align.py:251-252)
            self._active_items = [(n, p) for n, p in named.items()
                                  if n in self._engine.grads and n not in self._untouched]
            if self._exchange is not None:
                self._engine.gather = self._exchange
            if self._sync_group is not None:
                self._engine.enable_grad_sync(self._sync_group)
                self._active_items = [(n, p) for n, p in named.items()
                                      if n in self._engine.grads and n not in self._untouched]
        return self._engine

    def attach_exchange(self, exchange):
        """Multi-GPU embedding / LSE exchange (segclip_b200.p2p.EmbeddingExchange)."""
        self._exchange = exchange
        if self._engine is not None:
            self._engine.gather = exchange

    def enable_native_grad_sync(self, group):
        """Average gradients across `group` inside the native backward (bucketed NCCL all-reduce overlapped with the
        remaining backward kernels).  Use INSTEAD of DistributedDataParallel, before the first forward."""
        assert self._engine is None, "call before the first forward"
        self._sync_group = group

    def inject_noise(self, noise):
        """Replay explicit uniform draws {u1,u2,u3} instead of torch.rand (parity tests)."""
        self._noise = noise

    def force_assignment(self, forced):
        """Teacher-force the hard assignment {main: [B,L], mae: [B,L']}."""
        self._forced = forced

    # ---- the hot path ------------------------------------------------------------------------------
    def forward(self, input_ids, token_type_ids, attention_mask, image, image_seg=None):
        """modules/modeling.py:174-256.  token_type_ids / attention_mask are accepted and unused,
        exactly like the reference (padding is not masked, only causality; Appendix B.11)."""
        if not self.training:
            return None
        eng = self._get_engine()
        dev = eng.dev
        ids = torch.as_tensor(input_ids)
        ids = ids.view(-1, ids.shape[-1]).to(dev, non_blocking=True)
        img = torch.as_tensor(image)
        b, pair, ch, h, w = img.shape
        if img.dtype == torch.uint8:
            pass                                   # raw pixels: normalised on the device (CLIP mean / std), see _upload_u8
        elif img.device.type != "cpu":
            img = img[:, 0].float()                # device input: cast in place of the reference's .float() (modeling.py:182)
        elif img.dtype != torch.float32:
            img = img[:, 0].float()                # float64 host arrays of the reference loaders: cast once on the host
        seg = None
        if self.use_seglabel:
            seg = torch.as_tensor(image_seg)[:, 0].to(dev, non_blocking=True).reshape(b, -1).long()
        noise = self._noise
        if noise is None:
            c = eng
            noise = dict(u1=torch.rand(b, G, c.Lp, device=dev))
            if self.cfg["use_mae"]:
                noise["u2"] = torch.rand(b, c.Lp + 1, device=dev)
                noise["u3"] = torch.rand(b, G, c.Lm, device=dev)
        inputs = dict(ids=ids, image=img, seg=seg, norm=self.image_norm)
        params = [p for _, p in self._active_items]
        return _NativeStep.apply(self, b, inputs, noise, self._forced, *params)

    # ---- inference-mode getters (modules/modeling.py:258-372) ---------------------------------------
    def get_sequence_output(self, input_ids, token_type_ids, attention_mask, shaped=False, return_hidden=False, seq_model=None,
                            mask_ratio=0.):
        ids = torch.as_tensor(input_ids)
        ids = ids.view(-1, ids.shape[-1])
        bs = ids.size(0)
        out = self.clip.encode_text(ids, return_hidden=return_hidden, mask_ratio=mask_ratio)
        if isinstance(out, tuple):
            return tuple(t.float().view(bs, -1, t.size(-1)) for t in out)
        return out.float().view(bs, -1, out.size(-1))

    def get_visual_output(self, image, shaped=False, image_frame=-1, return_hidden=False, vis_model=None, mask_ratio=0.):
        image = torch.as_tensor(image)
        if shaped is False:
            b, pair, channel, h, w = image.shape
            image = image[:, 0].reshape(b, channel, h, w)
        bs = image.size(0)
        out = self.clip.encode_image(image, return_hidden=return_hidden, mask_ratio=mask_ratio)
        if isinstance(out, tuple):
            return tuple([t.float().view(bs, -1, t.size(-1)) for t in out[:2]] + [out[2]])
        return out.float().view(bs, -1, out.size(-1))

    def get_sequence_visual_output(self, input_ids, token_type_ids, attention_mask, image, shaped=False, image_frame=-1,
                                   return_hidden=False, seq_model=None, vis_model=None):
        seq = self.get_sequence_output(input_ids, token_type_ids, attention_mask, shaped=shaped, return_hidden=return_hidden)
        vis = self.get_visual_output(image, shaped=shaped, image_frame=image_frame, return_hidden=return_hidden)
        return seq, vis

    def _loose_similarity(self, sequence_output, visual_output, logit_scale=None):
        """Inference branch of modules/modeling.py:338-362: cosine logits scaled by min(exp(logit_scale), 100)."""
        if self.training:
            raise NotImplementedError("the training-mode similarity (all-gather InfoNCE) is fused into SegCLIP.forward()")
        dev = self._get_engine().dev
        t = sequence_output.squeeze(1).contiguous().float().to(dev)
        v = visual_output.squeeze(1).contiguous().float().to(dev)
        tn, vn = torch.empty_like(t), torch.empty_like(v)
        ti, vi = torch.empty(t.shape[0], device=dev), torch.empty(v.shape[0], device=dev)
        ops.l2norm_fwd_op(t, tn, ti)()
        ops.l2norm_fwd_op(v, vn, vi)()
        p = self.clip.logit_scale if logit_scale is None else logit_scale
        scale = min(math.exp(float(p)), 100.0)
        t2v = torch.empty(t.shape[0], v.shape[0], device=dev)
        ops.gemm_op(tn, vn, t2v, alpha=scale)()
        return t2v, t2v.T

    def get_similarity_logits(self, sequence_output, visual_output, attention_mask, shaped=False):
        t2v, v2t = self._loose_similarity(sequence_output, visual_output)
        return t2v, v2t, ()

    # ---- introspection used by tests -----------------------------------------------------------
    def debug_buffers(self, B):
        return self._engine.plan(B).bufs


def _load_clip_archive(name):
    files = {"ViT-B/32": "ViT-B-32.pt", "ViT-B/16": "ViT-B-16.pt", "ViT-L/14": "ViT-L-14.pt"}
    cands = []
    if name in files:
        here = os.path.dirname(os.path.abspath(__file__))
        cands = [os.path.join(here, files[name]), os.path.expanduser(os.path.join("~/.cache/clip", files[name]))]
    elif os.path.isfile(name):
        cands = [name]
    for path in cands:
        if os.path.isfile(path):
            try:
                return torch.jit.load(path, map_location="cpu").eval().state_dict()
            except RuntimeError:
                return torch.load(path, map_location="cpu")
    raise RuntimeError("CLIP weights for %r not found on disk (looked in %s); pass clip_state_dict= or a file path -- "
                       "this build never downloads" % (name, cands))
