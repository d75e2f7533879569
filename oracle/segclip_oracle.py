"""TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") of the SegCLIP training hot path.

A functional, fp32, plain-PyTorch-on-CPU restatement of what the reference computes in
``SegCLIP.forward`` (modules/modeling.py:174-256) and everything below it.  It is the
checker for the CUDA path (tests/, __graft_entry__.smoke) and the CPU baseline in
bench.py.  The product package never imports it.

Pinning: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (imported from
/root/reference through oracle/ref_harness.py) on seeded inputs and commits loss /
gradient fixtures under tests/golden/; ``tests/test_oracle_golden.py`` checks this file
against them, and ``tests/test_oracle_vs_reference.py`` re-checks against the live
reference whenever /root/reference is present.

Every function cites the reference lines it restates.  Parameters are a flat
``{state_dict key: tensor}`` mapping using the reference's own key names (SURVEY.md
Appendix A), so a reference ``state_dict()`` can be passed in unchanged.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

NUM_CENTERS = 8          # modules/module_seg_vit.py:349 (group_num=8)
GUMBEL_TAU = 0.9         # modules/module_seg_vit.py:305
DEC_HEADS = 8            # modules/modeling.py:147
DEC_DEPTH = 3            # modules/modeling.py:151


def vit_b16_config(**over):
    cfg = dict(vision_width=768, text_width=512, embed_dim=512, patch=16, grid=14,
               context=77, vocab=49408, text_layers=12, first_stage_layer=10,
               use_mae=False, use_kl=False)
    cfg.update(over)
    return cfg


def toy_config(**over):
    """Small model with the same structure (2 vision heads, 1 text head, 4x4 patches)."""
    cfg = dict(vision_width=128, text_width=64, embed_dim=64, patch=16, grid=4,
               context=16, vocab=512, text_layers=2, first_stage_layer=10,
               use_mae=False, use_kl=False)
    cfg.update(over)
    return cfg


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def _block_keys(prefix, d, g, p):
    """CLIP-style residual block parameters (modules/module_seg_vit.py:162-174)."""
    p[prefix + "attn.in_proj_weight"] = torch.randn(3 * d, d, generator=g) * d ** -0.5
    p[prefix + "attn.in_proj_bias"] = torch.randn(3 * d, generator=g) * 0.02
    p[prefix + "attn.out_proj.weight"] = torch.randn(d, d, generator=g) * 0.02
    p[prefix + "attn.out_proj.bias"] = torch.randn(d, generator=g) * 0.02
    for ln in ("ln_1", "ln_2"):
        p[prefix + ln + ".weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        p[prefix + ln + ".bias"] = 0.05 * torch.randn(d, generator=g)
    p[prefix + "mlp.c_fc.weight"] = torch.randn(4 * d, d, generator=g) * 0.02
    p[prefix + "mlp.c_fc.bias"] = torch.randn(4 * d, generator=g) * 0.02
    p[prefix + "mlp.c_proj.weight"] = torch.randn(d, 4 * d, generator=g) * 0.02
    p[prefix + "mlp.c_proj.bias"] = torch.randn(d, generator=g) * 0.02


def init_params(cfg, seed=0, center_std=1.0):
    """Deterministic synthetic parameters with the reference's names and shapes.

    Distributions follow the reference's initialisers in spirit (N(0,0.02) Linear weights,
    modules/util_module.py:70-85) but biases / LN affine terms are perturbed so that every
    parameter influences the loss (a stricter parity test than zeros/ones).  ``center_std``
    = 1 separates the semantic centers (avoids near-tie argmax at init).
    """
    g = torch.Generator().manual_seed(seed)
    vw, tw, e = cfg["vision_width"], cfg["text_width"], cfg["embed_dim"]
    ps, gr = cfg["patch"], cfg["grid"]
    fsl = cfg["first_stage_layer"]
    p = OrderedDict()
    p["clip.logit_scale"] = torch.tensor(math.log(1 / 0.07))
    p["clip.positional_embedding"] = torch.randn(cfg["context"], tw, generator=g) * 0.01
    p["clip.text_projection"] = torch.randn(tw, e, generator=g) * tw ** -0.5
    p["clip.token_embedding.weight"] = torch.randn(cfg["vocab"], tw, generator=g) * 0.02
    p["clip.ln_final.weight"] = 1.0 + 0.1 * torch.randn(tw, generator=g)
    p["clip.ln_final.bias"] = 0.05 * torch.randn(tw, generator=g)
    v = "clip.visual."
    p[v + "class_embedding"] = torch.randn(vw, generator=g) * vw ** -0.5
    p[v + "positional_embedding"] = torch.randn(gr * gr + 1, vw, generator=g) * vw ** -0.5
    p[v + "proj"] = torch.randn(vw, e, generator=g) * vw ** -0.5
    p[v + "conv1.weight"] = torch.randn(vw, 3, ps, ps, generator=g) * 0.02
    for ln in ("ln_pre", "ln_post"):
        p[v + ln + ".weight"] = 1.0 + 0.1 * torch.randn(vw, generator=g)
        p[v + ln + ".bias"] = 0.05 * torch.randn(vw, generator=g)
    t = v + "transformer."
    for i in range(fsl):
        _block_keys(f"{t}layers0.{i}.", vw, g, p)
    for i in range(12 - fsl):
        _block_keys(f"{t}layers2.{i}.", vw, g, p)
    for i in range(12 - fsl):
        _block_keys(f"{t}layers_mae2.{i}.", vw, g, p)
    s = t + "semantic_layer2."
    p[s + "semantic_center"] = torch.randn(NUM_CENTERS, vw, generator=g) * center_std
    for ln in ("norm", "cross_ln", "k_ln", "proj_o.ln"):
        p[s + ln + ".weight"] = 1.0 + 0.1 * torch.randn(vw, generator=g)
        p[s + ln + ".bias"] = 0.05 * torch.randn(vw, generator=g)
    for i in range(2):
        c = f"{s}cross_att.{i}."
        p[c + "attn.in_proj_weight"] = torch.randn(3 * vw, vw, generator=g) * vw ** -0.5
        p[c + "attn.in_proj_bias"] = torch.randn(3 * vw, generator=g) * 0.02
        p[c + "attn.out_proj.weight"] = torch.randn(vw, vw, generator=g) * 0.02
        p[c + "attn.out_proj.bias"] = torch.randn(vw, generator=g) * 0.02
        for ln in ("ln_x", "ln_k", "ln_2"):
            p[c + ln + ".weight"] = 1.0 + 0.1 * torch.randn(vw, generator=g)
            p[c + ln + ".bias"] = 0.05 * torch.randn(vw, generator=g)
        p[c + "mlp.c_fc.weight"] = torch.randn(4 * vw, vw, generator=g) * 0.02
        p[c + "mlp.c_fc.bias"] = torch.randn(4 * vw, generator=g) * 0.02
        p[c + "mlp.c_proj.weight"] = torch.randn(vw, 4 * vw, generator=g) * 0.02
        p[c + "mlp.c_proj.bias"] = torch.randn(vw, generator=g) * 0.02
    p[s + "k_conv.weight"] = torch.randn(vw, 64, 1, generator=g) * 0.125
    p[s + "v_conv.weight"] = torch.randn(vw, 64, 1, generator=g) * 0.125
    p[s + "proj_o.mlp.fc1.weight"] = torch.randn(4 * vw, vw, generator=g) * 0.02
    p[s + "proj_o.mlp.fc1.bias"] = torch.randn(4 * vw, generator=g) * 0.02
    p[s + "proj_o.mlp.fc2.weight"] = torch.randn(vw, 4 * vw, generator=g) * 0.02
    p[s + "proj_o.mlp.fc2.bias"] = torch.randn(vw, generator=g) * 0.02
    r = t + "reconstruct_layer2.rec_proj_a.a_fc."
    p[r + "weight"] = torch.randn(NUM_CENTERS, NUM_CENTERS, generator=g) * 0.3
    p[r + "bias"] = torch.randn(NUM_CENTERS, generator=g) * 0.02
    for i in range(cfg["text_layers"]):
        _block_keys(f"clip.transformer.resblocks.{i}.", tw, g, p)
    if cfg["use_mae"]:
        dd = vw // 2
        m = "vis_mae_decoder."
        p[m + "mask_token"] = torch.randn(1, 1, dd, generator=g) * 0.02
        p[m + "decoder_pos_embed"] = torch.from_numpy(
            sincos_2d_pos_embed(dd, gr, cls_token=True)).float().unsqueeze(0)
        p[m + "decoder_embed.weight"] = torch.randn(dd, vw, generator=g) * vw ** -0.5
        p[m + "decoder_embed.bias"] = torch.randn(dd, generator=g) * 0.02
        for i in range(DEC_DEPTH):
            b = f"{m}decoder_blocks.{i}."
            for ln in ("norm1", "norm2"):
                p[b + ln + ".weight"] = 1.0 + 0.1 * torch.randn(dd, generator=g)
                p[b + ln + ".bias"] = 0.05 * torch.randn(dd, generator=g)
            p[b + "attn.qkv.weight"] = torch.randn(3 * dd, dd, generator=g) * dd ** -0.5
            p[b + "attn.qkv.bias"] = torch.randn(3 * dd, generator=g) * 0.02
            p[b + "attn.proj.weight"] = torch.randn(dd, dd, generator=g) * 0.02
            p[b + "attn.proj.bias"] = torch.randn(dd, generator=g) * 0.02
            p[b + "mlp.fc1.weight"] = torch.randn(4 * dd, dd, generator=g) * 0.02
            p[b + "mlp.fc1.bias"] = torch.randn(4 * dd, generator=g) * 0.02
            p[b + "mlp.fc2.weight"] = torch.randn(dd, 4 * dd, generator=g) * 0.02
            p[b + "mlp.fc2.bias"] = torch.randn(dd, generator=g) * 0.02
        p[m + "decoder_norm.weight"] = 1.0 + 0.1 * torch.randn(dd, generator=g)
        p[m + "decoder_norm.bias"] = 0.05 * torch.randn(dd, generator=g)
        p[m + "decoder_pred.weight"] = torch.randn(3 * ps * ps, dd, generator=g) * dd ** -0.5
        p[m + "decoder_pred.bias"] = torch.randn(3 * ps * ps, generator=g) * 0.02
    return p


def sincos_2d_pos_embed(dim, grid, cls_token=True):
    """Fixed 2-D sin-cos table of the MAE decoder (modules/module_mae.py:63-108): the
    W-coordinate grid feeds the first half of the channels, the H-coordinate grid the
    second half; each half is [sin(pos*w_k) | cos(pos*w_k)], w_k = 10000^(-k/(dim/4))."""
    def one_dim(d, pos):
        omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=np.float64) / (d / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    gh = np.arange(grid, dtype=np.float32)
    gw = np.arange(grid, dtype=np.float32)
    mesh = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid, grid)
    emb = np.concatenate([one_dim(dim // 2, mesh[0]), one_dim(dim // 2, mesh[1])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, dim]), emb], axis=0)
    return emb


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def make_batch(cfg, batch, seed=0, rank=0):
    g = torch.Generator().manual_seed(seed + rank)
    res = cfg["patch"] * cfg["grid"]
    T, V = cfg["context"], cfg["vocab"]
    L = cfg["grid"] ** 2
    image = torch.randn(batch, 1, 3, res, res, generator=g)
    ids = torch.zeros(batch, 1, T, dtype=torch.int64)
    for b in range(batch):
        n = int(torch.randint(min(5, T - 3), T - 1, (1,), generator=g))
        ids[b, 0, 0] = V - 2                                   # SOT
        ids[b, 0, 1:1 + n] = torch.randint(1, V - 2, (n,), generator=g)
        ids[b, 0, 1 + n] = V - 1                               # EOT = unique arg-max
    keep = int((L + 1) * (1 - cfg.get("mae_vis_mask_ratio", 0.75))) - 1      # module_clip_util.py:98, CLS dropped
    noise = dict(u1=torch.rand(batch, NUM_CENTERS, L, generator=g),
                 u2=torch.rand(batch, L + 1, generator=g),
                 u3=torch.rand(batch, NUM_CENTERS, keep, generator=g))
    seg = torch.randint(0, 6, (batch, 1, cfg["grid"], cfg["grid"]), generator=g)
    return dict(input_ids=ids, attention_mask=(ids != 0).long(), image=image, image_seg=seg), noise


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def quick_gelu(x):
    """modules/module_clip_util.py:134-136"""
    return x * torch.sigmoid(1.702 * x)


def ln(x, p, prefix, eps=1e-5):
    """modules/module_clip_util.py:126-132 (fp32 LayerNorm)."""
    return F.layer_norm(x, (x.shape[-1],), p[prefix + ".weight"], p[prefix + ".bias"], eps)


def gumbel_from_uniform(u):
    """torch.distributions.Gumbel(0,1).sample() as a function of the underlying
    torch.rand draw u (modules/module_seg_vit.py:223-226): Uniform(tiny, 1-eps) base
    distribution followed by -log(-log(.))."""
    fi = torch.finfo(u.dtype)
    base = fi.tiny + u * ((1 - fi.eps) - fi.tiny)
    return -torch.log(-torch.log(base))


def attention_core(q, k, v, heads, mask=None):
    """softmax(q k^T / sqrt(hd) + mask) v with inputs in [N, L, D] (nn.MultiheadAttention
    semantics; q is scaled after projection)."""
    n, lq, d = q.shape
    lk = k.shape[1]
    hd = d // heads
    qh = q.view(n, lq, heads, hd).transpose(1, 2) * hd ** -0.5
    kh = k.view(n, lk, heads, hd).transpose(1, 2)
    vh = v.view(n, lk, heads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    o = torch.softmax(s, dim=-1) @ vh
    return o.transpose(1, 2).reshape(n, lq, d)


def self_attn_block(x, p, pre, heads, mask=None):
    """Pre-LN residual block, modules/module_seg_vit.py:191-196 and
    modules/module_clip_ttransformer.py:48-52 (identical math, NLD here)."""
    h = ln(x, p, pre + "ln_1")
    qkv = F.linear(h, p[pre + "attn.in_proj_weight"], p[pre + "attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    a = attention_core(q, k, v, heads, mask)
    x = x + F.linear(a, p[pre + "attn.out_proj.weight"], p[pre + "attn.out_proj.bias"])
    h = ln(x, p, pre + "ln_2")
    h = quick_gelu(F.linear(h, p[pre + "mlp.c_fc.weight"], p[pre + "mlp.c_fc.bias"]))
    return x + F.linear(h, p[pre + "mlp.c_proj.weight"], p[pre + "mlp.c_proj.bias"])


def cross_attn_block(q, kv, p, pre, heads, kv_layout):
    """modules/module_seg_vit.py:213-218.  q [B,G,D]; kv [B,S,D].

    kv_layout="torch18_flat" reproduces the pinned torch-1.8 behaviour: the
    projected K/V matrix [B*S, D] is re-read as [S, B, D], i.e. batch slot b' at key
    position s uses flat row s*B+b'.  "per_sample" is the intended semantics."""
    b, s, d = kv.shape
    w, bias = p[pre + "attn.in_proj_weight"], p[pre + "attn.in_proj_bias"]
    qn = ln(q, p, pre + "ln_x")
    kn = ln(kv, p, pre + "ln_k")
    qp = F.linear(qn, w[:d], bias[:d])
    kp = F.linear(kn, w[d:2 * d], bias[d:2 * d])
    vp = F.linear(kn, w[2 * d:], bias[2 * d:])
    if kv_layout == "torch18_flat":
        kp = kp.reshape(s, b, d).transpose(0, 1)
        vp = vp.reshape(s, b, d).transpose(0, 1)
    else:
        assert kv_layout == "per_sample"
    a = attention_core(qp, kp, vp, heads)
    q = q + F.linear(a, p[pre + "attn.out_proj.weight"], p[pre + "attn.out_proj.bias"])
    h = ln(q, p, pre + "ln_2")
    h = quick_gelu(F.linear(h, p[pre + "mlp.c_fc.weight"], p[pre + "mlp.c_fc.bias"]))
    return q + F.linear(h, p[pre + "mlp.c_proj.weight"], p[pre + "mlp.c_proj.bias"])


def grouped_conv1x1(x, w, groups):
    """nn.Conv1d(C, C, 1, groups=heads, bias=False) on [B,L,C] (module_seg_vit.py:266-269):
    block-diagonal linear with [C/groups x C/groups] blocks."""
    b, l, c = x.shape
    cg = c // groups
    xg = x.view(b, l, groups, cg)
    wg = w.view(groups, cg, cg)              # [g, out, in]
    return torch.einsum("blgi,goi->blgo", xg, wg).reshape(b, l, c)


def semantic_learner(x, p, pre, heads, u, kv_layout, forced_idx=None, training=True):
    """SemanticLearnerModule.forward, modules/module_seg_vit.py:277-314 (training mode).

    x [B,L,D]; u [B,G,L] uniform draw behind the Gumbel noise.  Returns
    (outputs [B,G,D], hard_attn [B,G,L] (straight-through), soft_attn, q_feat, idx [B,L])."""
    b, l, d = x.shape
    xin = ln(x, p, pre + "norm")
    q = p[pre + "semantic_center"].unsqueeze(0).expand(b, -1, -1)
    for i in range(2):
        kv = torch.cat([q, x], dim=1)
        q = cross_attn_block(q, kv, p, f"{pre}cross_att.{i}.", heads, kv_layout)
    q = ln(q, p, pre + "cross_ln")
    k = ln(grouped_conv1x1(xin, p[pre + "k_conv.weight"], heads), p, pre + "k_ln")
    v = grouped_conv1x1(xin, p[pre + "v_conv.weight"], heads)
    attn = torch.einsum("bgc,blc->bgl", q, k)
    if training:
        y = torch.softmax((attn + gumbel_from_uniform(u)) / GUMBEL_TAU, dim=1)
    else:                                                            # module_seg_vit.py:230-231: plain softmax, no noise, no tau
        y = torch.softmax(attn, dim=1)
    idx = y.argmax(dim=1) if forced_idx is None else forced_idx
    hard = torch.zeros_like(y).scatter_(1, idx.unsqueeze(1), 1.0)
    hard = hard - y.detach() + y                                     # :237 straight-through
    soft = torch.softmax(attn, dim=1)
    out = torch.einsum("bgl,blc->bgc", hard, v)
    out = out / torch.clamp_min(hard.sum(-1, keepdim=True), 1.0)
    h = ln(q + out, p, pre + "proj_o.ln")
    h = F.gelu(F.linear(h, p[pre + "proj_o.mlp.fc1.weight"], p[pre + "proj_o.mlp.fc1.bias"]))
    h = F.linear(h, p[pre + "proj_o.mlp.fc2.weight"], p[pre + "proj_o.mlp.fc2.bias"])
    return quick_gelu(h), hard, soft, q, idx


def reconstruct_layer(sx, hard, p, pre):
    """ReconstructLayer.forward, modules/module_seg_vit.py:333-345."""
    a = F.linear(hard.transpose(1, 2), p[pre + "rec_proj_a.a_fc.weight"], p[pre + "rec_proj_a.a_fc.bias"])
    return quick_gelu(torch.einsum("bdh,bmd->bmh", sx, a))


def random_masking_keep_cls(x, u2, mask_ratio=0.75):
    """modules/module_clip_util.py:91-124 with keep_cls=True; u2 is the torch.rand draw."""
    n, l, d = x.shape
    keep = int(l * (1 - mask_ratio))
    noise = u2.clone()
    noise[:, 0] = -1.0
    ids_shuffle = torch.argsort(noise, dim=1)
    ids_restore = torch.argsort(ids_shuffle, dim=1)
    ids_keep = ids_shuffle[:, :keep]
    xm = torch.gather(x, 1, ids_keep.unsqueeze(-1).expand(-1, -1, d))
    mask = torch.ones(n, l)
    mask[:, :keep] = 0
    mask = torch.gather(mask, 1, ids_restore)
    return xm, mask, ids_restore, ids_keep


def eval_pos_embed(pos, h_, w_):
    """VisualTransformer.get_pos_embed in eval mode (modules/module_clip_vtransformer.py:35-53): the patch part of the table
    is bicubically interpolated (align_corners=False) to the h_ x w_ grid of the input; the CLS row is kept."""
    pos_cls, pos_patch = pos[:1], pos[1:]
    n, dim = pos_patch.shape
    if h_ * w_ == n and h_ == w_:
        return pos
    g = int(math.sqrt(n))
    r = F.interpolate(pos_patch.reshape(1, g, g, dim).permute(0, 3, 1, 2), size=(h_, w_), mode="bicubic", align_corners=False)
    return torch.cat([pos_cls, r.permute(0, 2, 3, 1).reshape(-1, dim)], dim=0)


def patch_embed(image, p, cfg, training=True):
    """VisualTransformer.forward up to ln_pre, modules/module_clip_vtransformer.py:55-65 (training: the raw table, :36-37)."""
    v = "clip.visual."
    x = F.conv2d(image, p[v + "conv1.weight"], stride=cfg["patch"])
    h_, w_ = x.shape[-2:]
    x = x.flatten(2).transpose(1, 2)
    cls = p[v + "class_embedding"].expand(x.shape[0], 1, -1)
    pos = p[v + "positional_embedding"] if training else eval_pos_embed(p[v + "positional_embedding"], h_, w_)
    x = torch.cat([cls, x], dim=1) + pos
    return ln(x, p, v + "ln_pre")


def encode_image(image, p, cfg, u1, kv_layout, forced_idx=None, training=True, forced_pool=None):
    """Main visual pass: encode_image (modules/module_clip.py:81-103) + SegViT main branch
    (modules/module_seg_vit.py:434-448).  Returns (emb [B,E], aux).  ``forced_pool`` [B,D] teacher-forces the
    per-channel arg-max of the centre max-pooling (reduced-precision comparisons only)."""
    v, t = "clip.visual.", "clip.visual.transformer."
    heads = cfg["vision_width"] // 64
    x = patch_embed(image, p, cfg, training)[:, 1:]            # CLS dropped (:419)
    for i in range(cfg["first_stage_layer"]):
        x = self_attn_block(x, p, f"{t}layers0.{i}.", heads)
    sx, hard, soft, _, idx = semantic_learner(x, p, t + "semantic_layer2.", heads, u1, kv_layout, forced_idx, training)
    c = sx
    for i in range(12 - cfg["first_stage_layer"]):
        c = self_attn_block(c, p, f"{t}layers2.{i}.", heads)
    cls, pool_arg = c.max(dim=1, keepdim=True)
    if forced_pool is not None:
        pool_arg = forced_pool.unsqueeze(1)
        cls = c.gather(1, pool_arg)
    hid = ln(torch.cat([cls, c], dim=1), p, v + "ln_post") @ p[v + "proj"]
    return hid[:, 0], dict(hard_attn=hard, soft_attn=soft, assign=idx, patches=x, centers=c, hidden=hid,
                           pool_arg=pool_arg[:, 0])


def encode_image_mae(image, p, cfg, u2, u3, kv_layout, forced_idx=None):
    """Masked visual pass (modules/modeling.py:238-245, module_seg_vit.py:423-433): returns the
    decoder input [B, keep, D] (mean-CLS prepended), the mask and ids_restore."""
    t = "clip.visual.transformer."
    heads = cfg["vision_width"] // 64
    x = patch_embed(image, p, cfg)
    x, mask, ids_restore, _ = random_masking_keep_cls(x, u2, cfg.get("mae_vis_mask_ratio", 0.75))
    x = x[:, 1:]
    for i in range(cfg["first_stage_layer"]):
        x = self_attn_block(x, p, f"{t}layers0.{i}.", heads)
    sx, hard, _, _, idx = semantic_learner(x, p, t + "semantic_layer2.", heads, u3, kv_layout, forced_idx)
    x = reconstruct_layer(sx, hard, p, t + "reconstruct_layer2.")
    for i in range(12 - cfg["first_stage_layer"]):
        x = self_attn_block(x, p, f"{t}layers_mae2.{i}.", heads)
    x = torch.cat([x.mean(dim=1, keepdim=True), x], dim=1)
    return x, mask, ids_restore, idx


def patchify(img, ps):
    """modules/module_mae.py:18-29"""
    n, _, hh, _ = img.shape
    h = hh // ps
    x = img.reshape(n, 3, h, ps, h, ps)
    return torch.einsum("nchpwq->nhwpqc", x).reshape(n, h * h, ps * ps * 3)


def mae_decoder_loss(image, hid, mask, ids_restore, p, cfg):
    """MAEDecoder.forward_vis, modules/module_mae.py:304-330 (timm Block: LN eps 1e-6, qkv bias,
    8 heads, erf-GELU)."""
    m = "vis_mae_decoder."
    x = F.linear(hid, p[m + "decoder_embed.weight"], p[m + "decoder_embed.bias"])
    n, keep, dd = x.shape
    ltot = ids_restore.shape[1]
    x = torch.cat([x, p[m + "mask_token"].expand(n, ltot - keep, -1)], dim=1)
    x = torch.gather(x, 1, ids_restore.unsqueeze(-1).expand(-1, -1, dd))
    x = x + p[m + "decoder_pos_embed"]
    for i in range(DEC_DEPTH):
        b = f"{m}decoder_blocks.{i}."
        h = ln(x, p, b + "norm1", 1e-6)
        q, k, v = F.linear(h, p[b + "attn.qkv.weight"], p[b + "attn.qkv.bias"]).chunk(3, dim=-1)
        a = attention_core(q, k, v, DEC_HEADS)
        x = x + F.linear(a, p[b + "attn.proj.weight"], p[b + "attn.proj.bias"])
        h = ln(x, p, b + "norm2", 1e-6)
        h = F.gelu(F.linear(h, p[b + "mlp.fc1.weight"], p[b + "mlp.fc1.bias"]))
        x = x + F.linear(h, p[b + "mlp.fc2.weight"], p[b + "mlp.fc2.bias"])
    x = ln(x, p, m + "decoder_norm", 1e-6)
    pred = F.linear(x, p[m + "decoder_pred.weight"], p[m + "decoder_pred.bias"])[:, 1:]
    per_patch = ((pred - patchify(image, cfg["patch"])) ** 2).mean(dim=-1)
    mk = mask[:, 1:]
    return (per_patch * mk).sum() / mk.sum()


def encode_text(ids, p, cfg, return_hidden=False):
    """CLIP.encode_text, modules/module_clip.py:105-143 (causal mask, no padding mask)."""
    heads = cfg["text_width"] // 64
    t = ids.shape[1]
    x = p["clip.token_embedding.weight"][ids] + p["clip.positional_embedding"][:t]
    mask = torch.full((t, t), float("-inf")).triu_(1)          # module_clip_util.py:199-205
    for i in range(cfg["text_layers"]):
        x = self_attn_block(x, p, f"clip.transformer.resblocks.{i}.", heads, mask)
    hid = ln(x, p, "clip.ln_final") @ p["clip.text_projection"]
    out = hid[torch.arange(ids.shape[0]), ids.argmax(dim=-1)]
    return (out, hid) if return_hidden else out


def superpixel_kl(hard, seg):
    """modules/modeling.py:212-224.  hard [B,G,L] straight-through assignment, seg [B,g,g]."""
    a = hard.permute(0, 2, 1)
    b = a.shape[0]
    s = seg.reshape(b, -1)
    same = (s.unsqueeze(-1) == s.unsqueeze(-2)).to(a.dtype)
    mean = torch.einsum("bgl,blc->bgc", same, a) / torch.clamp_min(same.sum(-1, keepdim=True), 1.0)
    coef = float(a.numel())
    k1 = F.kl_div(F.log_softmax(a, -1), F.softmax(mean, -1), reduction="sum") / coef
    k2 = F.kl_div(F.log_softmax(mean, -1), F.softmax(a, -1), reduction="sum") / coef
    return (k1 + k2) / 2.0


def contrastive_loss(t_loc, v_loc, t_all, v_all, logit_scale_param, rank):
    """_loose_similarity + the two cross-entropies, modules/modeling.py:338-357,204-209.
    *_loc are this rank's L2-normalised rows, *_all the gathered [B*W, E] matrices."""
    scale = torch.clamp(logit_scale_param.exp(), max=100)
    t2v = scale * t_loc @ v_all.t()
    v2t = scale * v_loc @ t_all.t()
    b = t_loc.shape[0]
    labels = torch.arange(b) + b * rank
    return (F.cross_entropy(t2v, labels) + F.cross_entropy(v2t, labels)) / 2.0


def encode_image_eval(image, p, cfg, kv_layout="torch18_flat"):
    """Inference-mode CLIP.encode_image(image, return_hidden=True) (modules/module_clip.py:89-103 with the eval branch of
    gumbel_softmax): returns (x [B,E], hidden [B,9,E], mid_states) like the reference."""
    x, aux = encode_image(image, p, cfg, None, kv_layout, None, training=False)
    mid = {"hidden": aux["patches"], "attns": [{"soft_attn": aux["soft_attn"], "hard_attn": aux["hard_attn"]}]}
    return x, aux["hidden"], mid


def l2_normalize(x):
    return x / x.norm(dim=-1, keepdim=True)


def rank_forward(p, batch, noise, cfg, kv_layout="torch18_flat", forced=None):
    """Everything of SegCLIP.forward that is rank-local: returns normalised (t, v) embeddings and
    the auxiliary (KL + MAE) loss of this rank.  ``forced`` = {"main": idx [B,L], "mae": idx
    [B,keep-1]} teacher-forces the hard assignment; "pool" [B,D] the max-pooling arg-max."""
    forced = forced or {}
    ids = batch["input_ids"].view(-1, batch["input_ids"].shape[-1])
    image = batch["image"].float()[:, 0]
    t = encode_text(ids, p, cfg)
    v, aux = encode_image(image, p, cfg, noise["u1"], kv_layout, forced.get("main"), forced_pool=forced.get("pool"))
    extra = torch.zeros(())
    info = dict(assign_main=aux["assign"], hard_attn=aux["hard_attn"], soft_attn=aux["soft_attn"],
                t_raw=t, v_raw=v, pool_arg=aux["pool_arg"])
    if cfg["use_kl"]:
        info["kl"] = superpixel_kl(aux["hard_attn"], batch["image_seg"][:, 0])
        extra = extra + info["kl"]
    if cfg["use_mae"]:
        hid, mask, ids_restore, idx = encode_image_mae(image, p, cfg, noise["u2"], noise["u3"],
                                                       kv_layout, forced.get("mae"))
        info["mae"] = mae_decoder_loss(image, hid, mask, ids_restore, p, cfg)
        info["assign_mae"] = idx
        info["mae_mask"], info["mae_ids_restore"] = mask, ids_restore
        extra = extra + info["mae"]
    return l2_normalize(t), l2_normalize(v), extra, info


def forward(p, batch, noise, cfg, kv_layout="torch18_flat", forced=None):
    """World-size-1 training loss of SegCLIP.forward (modules/modeling.py:174-254)."""
    t, v, extra, info = rank_forward(p, batch, noise, cfg, kv_layout, forced)
    info["contrastive"] = contrastive_loss(t, v, t, v, p["clip.logit_scale"], 0)
    return info["contrastive"] + extra, info


def forward_multi_rank(p, batches, noises, cfg, kv_layout="torch18_flat", forced=None):
    """W simulated ranks in one process.  Returns per-rank losses; ``mean(losses).backward()``
    yields exactly the gradients DDP leaves in ``.grad`` (diffdist sums embedding gradients over
    ranks, modules/util_module.py:180-190; DDP then averages parameter gradients)."""
    outs = [rank_forward(p, b, n, cfg, kv_layout, (forced or [None] * len(batches))[i])
            for i, (b, n) in enumerate(zip(batches, noises))]
    t_all = torch.cat([o[0] for o in outs])
    v_all = torch.cat([o[1] for o in outs])
    losses = [contrastive_loss(o[0], o[1], t_all, v_all, p["clip.logit_scale"], r) + o[2]
              for r, o in enumerate(outs)]
    return losses, [o[3] for o in outs]


def loss_and_grads(p, batch, noise, cfg, kv_layout="torch18_flat", forced=None, frozen=()):
    """Convenience: fp32 loss + {name: grad} by autograd."""
    q = OrderedDict((k, v.detach().clone().requires_grad_(v.is_floating_point() and k not in frozen))
                    for k, v in p.items())
    loss, info = forward(q, batch, noise, cfg, kv_layout, forced)
    loss.backward()
    grads = {k: v.grad for k, v in q.items() if v.grad is not None}
    return loss.detach(), grads, info
