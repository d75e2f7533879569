"""Geometry presets and the shape-only CLIP state dict used for random-initialised models.

The reference derives every size of the model from the SHAPES of a CLIP checkpoint's tensors
(modules/modeling.py:88-101: vision width from `visual.conv1.weight`, grid from
`visual.positional_embedding`, embed dim from `text_projection`, context / vocab / text width from
`positional_embedding` / `token_embedding.weight` / `ln_final.weight`, text depth from the number of
`transformer.resblocks.*` keys).  With no network there is no checkpoint, so benchmarks and tools build
the module from a dict of zero tensors of the right shapes; `SegCLIP.__init__` then random-initialises
exactly as the reference does.
"""
import torch


def vit_b16(**over):
    """ViT-B/16 + 12-layer text-77 tower (BASELINE configs[0..3])."""
    cfg = dict(vision_width=768, text_width=512, embed_dim=512, patch=16, grid=14, context=77, vocab=49408,
               text_layers=12, first_stage_layer=10, use_mae=False, use_kl=False)
    cfg.update(over)
    return cfg


def vit_l14(**over):
    """"ViT-L/14" as the reference builds it from an L/14-shaped checkpoint (SegViT hard-codes
    10+2 layers; width 1024 / 16 heads / patch 14 / 16x16 grid, text width 768, embed 768)."""
    return vit_b16(vision_width=1024, text_width=768, embed_dim=768, patch=14, grid=16, **over)


def shape_state_dict(cfg):
    """Zero tensors carrying only the shapes `SegCLIP.__init__` reads (modules/modeling.py:88-101)."""
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
        sd["visual.transformer.resblocks.%d.attn.in_proj_weight" % i] = torch.zeros(1)
    for i in range(cfg["text_layers"]):
        sd["transformer.resblocks.%d.ln_1.weight" % i] = torch.zeros(1)
    return sd
