"""Zero-shot segmentation inference over the B200 engine: the consumer of eval-mode
``clip.encode_image(img, return_hidden=True)`` that the reference implements as an mmseg ``EncoderDecoder`` subclass,
``ViTSegInference`` (seg_segmentation/evaluation/vit_seg.py:118-256), plus mmseg's whole / slide-window test-time drivers
(``EncoderDecoder.whole_inference`` / ``slide_inference``), without the mmcv / mmseg dependency.

Every encoder launch goes through the native engine (patch embedding, 10+2 blocks, centre aggregation with the plain softmax
assignment, ``soft_attn``); what is left here is the reference's own per-image glue on tiny tensors -- 8 centres x N classes
affinities, a bilinear up-sampling of the 8 attention maps, an arg-max -- which the reference also runs as plain PyTorch ops.
Slide-window mode stacks all crops of an image into ONE batched encoder call instead of mmseg's crop-by-crop loop (batch 1).
"""
import math

import torch
import torch.nn.functional as F


def resize_attn_map(attn, h, w, align_corners=False):
    """[B, nH, H'*W', G] -> [B, nH, h, w, G] (vit_seg.py:31-59): the token grid is recovered from the image aspect ratio."""
    n = attn.shape[2]
    scale = (h * w // n) ** 0.5
    if h > w:
        wf = w // int(round(scale))
        hf = n // wf
    else:
        hf = h // int(round(scale))
        wf = n // hf
    assert n == hf * wf, "%d tokens do not form a %d x %d grid (h=%d w=%d)" % (n, hf, wf, h, w)
    bs, nh, _, g = attn.shape
    x = attn.reshape(bs * nh, hf, wf, g).permute(0, 3, 1, 2)
    x = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=align_corners)
    return x.permute(0, 2, 3, 1).reshape(bs, nh, h, w, g)


def build_text_embedding(model, class_tokens):
    """[N, T, L] token ids of N classes x T prompt templates -> [N, C] normalised class embeddings
    (seg_segmentation/evaluation/builder.py:46-66)."""
    n, t, l = class_tokens.shape
    with torch.no_grad():
        e = model.clip.encode_text(class_tokens.reshape(n * t, l)).float()
    e = e.reshape(n, t, -1).mean(dim=1)
    return e / e.norm(dim=-1, keepdim=True)


class ViTSegInference(torch.nn.Module):
    """vit_seg.py:118-256.  ``test_cfg``: mode 'whole' | 'slide', bg_thresh, and for 'slide' crop_size / stride (pixels)."""

    def __init__(self, model, text_embedding, with_bg, test_cfg=None):
        super().__init__()
        cfg = dict(mode="whole", bg_thresh=0.95)
        cfg.update(test_cfg or {})
        self.test_cfg = cfg
        self.model = model
        self.register_buffer("text_embedding", text_embedding)
        self.with_bg = with_bg
        self.bg_thresh = cfg["bg_thresh"]
        self.num_classes = len(text_embedding) + (1 if with_bg else 0)
        self.align_corners = False

    # ---- encoder ---------------------------------------------------------------------------------
    def _encode(self, img):
        with torch.no_grad():
            x, hidden, mid = self.model.clip.encode_image(img, return_hidden=True)
        return x.float(), hidden.float(), mid

    def get_attn_maps(self, img, return_onehot=False, rescale=False, _mid=None):
        """list of [B, H, W, G] attention maps (vit_seg.py:144-200); SegCLIP has one grouping stage."""
        mid = _mid if _mid is not None else self._encode(img)[2]
        maps, prev = [], None
        for att in mid["attns"]:
            a = att["soft_attn"].float().unsqueeze(1).transpose(2, 3)          # [B, 1, HW, G]
            prev = a if prev is None else prev @ a
            maps.append(resize_attn_map(prev, *img.shape[-2:]))
        out = []
        for m in maps:
            assert m.shape[1] == 1
            m = m.squeeze(1)
            if rescale:
                m = F.interpolate(m.permute(0, 3, 1, 2), size=img.shape[2:], mode="bilinear",
                                  align_corners=self.align_corners).permute(0, 2, 3, 1)
            if return_onehot:
                m = F.one_hot(m.argmax(dim=-1), num_classes=m.shape[-1]).to(m.dtype)
            out.append(m)
        return out

    def encode_decode(self, img, img_metas=None):
        """[B, 3, H, W] -> per-pixel class logits [B, num_classes, H, W] (vit_seg.py:202-256; the reference asserts B == 1 and
        runs the encoder twice per image -- here one batched encoder call serves both the attention maps and the features)."""
        x, hidden, mid = self._encode(img)
        attn = self.get_attn_maps(img, rescale=True, _mid=mid)[-1]             # [B, H, W, G]
        tokens = F.normalize(hidden[:, 1:, :], dim=-1)                         # [B, G, C] grouped image tokens
        avg = F.normalize(x, dim=-1)                                           # [B, C]
        onehot = F.one_hot(attn.argmax(dim=-1), num_classes=attn.shape[-1]).to(attn.dtype)
        text = self.text_embedding.to(tokens.dtype)
        nfg = text.shape[0]
        off = 1 if self.with_bg else 0
        scale = torch.clamp(self.model.clip.logit_scale.detach().float().exp(), max=100)
        group_aff = (tokens @ text.T) * scale                                  # [B, G, N]
        pre = F.softmax(group_aff, dim=-1)
        avg_aff = F.softmax((avg @ text.T) * scale, dim=-1)                    # [B, N]
        top = avg_aff.topk(dim=-1, k=min(5, nfg))
        mask = torch.zeros_like(avg_aff).scatter_add_(-1, top.indices, torch.ones_like(top.values))
        group_aff = group_aff.masked_fill(~mask.bool().unsqueeze(1), float("-inf"))
        group_aff = F.softmax(group_aff, dim=-1) * pre
        B, H, W, _ = attn.shape
        logits = torch.zeros(B, nfg + off, H, W, device=img.device, dtype=attn.dtype)
        per_pixel = onehot @ group_aff.unsqueeze(1)                            # [B, H, W, N]
        logits[:, off:] = per_pixel.permute(0, 3, 1, 2)
        if self.with_bg:
            for i in range(B):
                thr = min(self.bg_thresh, float(group_aff[i].max()))
                logits[i, 0][per_pixel[i].max(dim=-1).values < thr] = 1
        return logits

    # ---- mmseg EncoderDecoder test-time drivers ---------------------------------------------------
    def whole_inference(self, img, rescale=False, ori_shape=None):
        logit = self.encode_decode(img)
        if rescale and ori_shape is not None:
            logit = F.interpolate(logit, size=ori_shape, mode="bilinear", align_corners=self.align_corners)
        return logit

    def slide_inference(self, img, rescale=False, ori_shape=None, max_batch=64):
        """Overlapping windows of crop_size with the given stride, logits averaged where windows overlap (mmseg
        EncoderDecoder.slide_inference).  All windows of the image are encoded in batched calls."""
        hs, ws = _pair(self.test_cfg["stride"])
        hc, wc = _pair(self.test_cfg["crop_size"])
        B, _, H, W = img.shape
        hg, wg = max(H - hc + hs - 1, 0) // hs + 1, max(W - wc + ws - 1, 0) // ws + 1
        boxes = []
        for hi in range(hg):
            for wi in range(wg):
                y2, x2 = min(hi * hs + hc, H), min(wi * ws + wc, W)
                boxes.append((max(y2 - hc, 0), y2, max(x2 - wc, 0), x2))
        preds = img.new_zeros((B, self.num_classes, H, W), dtype=torch.float32)
        count = img.new_zeros((B, 1, H, W), dtype=torch.float32)
        for s in range(0, len(boxes), max(max_batch // B, 1)):
            chunk = boxes[s:s + max(max_batch // B, 1)]
            crops = torch.cat([img[:, :, y1:y2, x1:x2] for (y1, y2, x1, x2) in chunk], dim=0)
            logit = self.encode_decode(crops).float()
            for k, (y1, y2, x1, x2) in enumerate(chunk):
                preds[:, :, y1:y2, x1:x2] += logit[k * B:(k + 1) * B]
                count[:, :, y1:y2, x1:x2] += 1
        assert int((count == 0).sum()) == 0
        preds = preds / count
        if rescale and ori_shape is not None:
            preds = F.interpolate(preds, size=ori_shape, mode="bilinear", align_corners=self.align_corners)
        return preds

    def inference(self, img, rescale=False, ori_shape=None):
        assert self.test_cfg["mode"] in ("slide", "whole")
        f = self.slide_inference if self.test_cfg["mode"] == "slide" else self.whole_inference
        return F.softmax(f(img, rescale, ori_shape), dim=1)

    def simple_test(self, img, rescale=True, ori_shape=None):
        """-> [B, H, W] int64 label map (mmseg EncoderDecoder.simple_test)."""
        return self.inference(img, rescale, ori_shape).argmax(dim=1)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)
