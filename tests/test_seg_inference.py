"""Zero-shot segmentation inference: inference at the reference's other input sizes (bicubic positional
table), the ViTSegInference consumer and the whole / slide-window drivers."""
import pytest
import torch
import torch.nn.functional as F


class _StubClip(torch.nn.Module):
    """CPU stand-in with the encode_image contract (x, hidden [B, 9, C], mid_states): deterministic functions of the pixels."""

    def __init__(self, patch=16, G=8, C=12):
        super().__init__()
        self.patch, self.G, self.C = patch, G, C
        self.logit_scale = torch.nn.Parameter(torch.tensor(2.0))
        g = torch.Generator().manual_seed(0)
        self.w = torch.randn(3, C, generator=g)
        self.calls = []

    def encode_image(self, img, return_hidden=False):
        self.calls.append(tuple(img.shape))
        B, _, H, W = img.shape
        gh, gw = H // self.patch, W // self.patch
        pooled = F.avg_pool2d(img, self.patch)                                   # [B, 3, gh, gw]
        tok = pooled.flatten(2).transpose(1, 2)                                   # [B, L, 3]
        soft = torch.softmax(torch.stack([tok[..., 0] * (k + 1) + tok[..., 1] * (self.G - k) for k in range(self.G)], 1), dim=1)
        centers = torch.einsum("bgl,blc->bgc", soft, tok) @ self.w                # [B, G, C]
        cls = centers.max(dim=1, keepdim=True)[0]
        hidden = torch.cat([cls, centers], 1)
        return hidden[:, 0], hidden, {"hidden": tok, "attns": [{"soft_attn": soft, "hard_attn": soft}]}


class _StubModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.clip = _StubClip()


def _seg(mode="whole", with_bg=True, **kw):
    from segclip_b200.seg_inference import ViTSegInference
    text = F.normalize(torch.randn(6, 12, generator=torch.Generator().manual_seed(1)), dim=-1)
    cfg = dict(mode=mode, bg_thresh=0.4)
    cfg.update(kw)
    return ViTSegInference(_StubModel(), text, with_bg, cfg)


def test_encode_decode_batched_equals_per_image_and_is_a_labelling():
    seg = _seg()
    img = torch.randn(3, 3, 64, 96, generator=torch.Generator().manual_seed(2))
    full = seg.encode_decode(img)
    assert full.shape == (3, 7, 64, 96)
    for i in range(3):
        one = seg.encode_decode(img[i:i + 1])
        assert torch.allclose(one[0], full[i], atol=1e-6)
    # every foreground pixel's logits are the affinities of exactly one group (one-hot attention map)
    assert float(full[:, 1:].sum(1).max()) <= 1.0 + 1e-5 and float(full.min()) >= 0.0


def test_slide_inference_full_window_equals_whole_and_windows_average():
    img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    whole = _seg("whole").whole_inference(img)
    slide = _seg("slide", crop_size=64, stride=64).slide_inference(img)
    assert torch.allclose(whole, slide, atol=1e-6)
    # overlapping windows: every pixel is the mean of the windows that cover it; all windows of an image go out in one call
    seg = _seg("slide", crop_size=32, stride=16)
    out = seg.slide_inference(img)
    assert seg.model.clip.calls == [(2 * 9, 3, 32, 32)]
    ref = torch.zeros_like(out)
    cnt = torch.zeros(2, 1, 64, 64)
    one = _seg("whole")
    for y in range(0, 33, 16):
        for x in range(0, 33, 16):
            ref[:, :, y:y + 32, x:x + 32] += one.encode_decode(img[:, :, y:y + 32, x:x + 32])
            cnt[:, :, y:y + 32, x:x + 32] += 1
    assert torch.allclose(out, ref / cnt, atol=1e-6)
    lab = seg.simple_test(img, rescale=True, ori_shape=(100, 80))
    assert lab.shape == (2, 100, 80) and lab.dtype == torch.int64 and int(lab.max()) < 7


def test_ragged_image_windows_are_shifted_inside():
    """Image size not a multiple of the stride: the last windows are shifted back inside the image (mmseg semantics)."""
    seg = _seg("slide", crop_size=32, stride=24)
    img = torch.randn(1, 3, 48, 80, generator=torch.Generator().manual_seed(4))
    out = seg.slide_inference(img)
    assert out.shape == (1, 7, 48, 80) and bool(torch.isfinite(out).all())


@pytest.mark.gpu
@pytest.mark.parametrize("src,dst", [((14, 14), (28, 28)), ((4, 4), (8, 8)), ((4, 4), (4, 16)), ((16, 16), (32, 32)), ((7, 7), (5, 9))])
def test_bicubic_resize_kernel_matches_torch(src, dst):
    from segclip_b200 import ops
    D = 136
    t = torch.randn(src[0] * src[1], D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    out = torch.empty(dst[0] * dst[1], D, device="cuda")
    ops.bicubic_resize_op(t, out, src, dst)()
    ref = F.interpolate(t.reshape(1, src[0], src[1], D).permute(0, 3, 1, 2), size=dst, mode="bicubic", align_corners=False)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, D)
    assert float((out - ref).abs().max()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(128, 128), (64, 256), (256, 64)])
def test_eval_encode_image_at_other_input_sizes_matches_oracle(hw):
    """fp32 eval-mode encode_image at the reference's other supported sizes (4 n patch tokens) against the oracle, which is
    pinned to the unmodified reference at the same sizes (tests/test_oracle_vs_reference.py)."""
    from oracle import segclip_oracle as so
    from tools.e2e_report import build_model
    cfg = so.toy_config()
    params = so.init_params(cfg, seed=43)
    img = torch.randn(2, 3, hw[0], hw[1], generator=torch.Generator().manual_seed(44))
    ox, ohid, omid = so.encode_image_eval(img, params, cfg)
    model = build_model(cfg, params, "fp32", "torch18_flat").eval()
    with torch.no_grad():
        x, hid, mid = model.clip.encode_image(img, return_hidden=True)
        x0 = model.clip.encode_image(img[:, :, :64, :64])              # the training size still works next to it
    def rel(a, b):
        return float((a.float().cpu() - b).abs().max() / (b.abs().max() + 1e-9))
    assert rel(x, ox) < 1e-4 and rel(hid, ohid) < 1e-4 and rel(mid["hidden"], omid["hidden"]) < 1e-4
    assert rel(mid["attns"][0]["soft_attn"], omid["attns"][0]["soft_attn"]) < 1e-4
    assert x0.shape == x.shape
    with pytest.raises(Exception):
        model.clip.encode_image(torch.randn(1, 3, 96, 96))             # 36 tokens: neither n nor 4 n (module_seg_vit.py:423)


@pytest.mark.gpu
def test_seg_inference_end_to_end_on_the_engine():
    """ViTSegInference over the real engine: slide-window mode on a 2x-size image, batched crops, labels in range; whole mode
    at 2x resolution (bicubic positional table) agrees with slide mode using one full-size window."""
    from oracle import segclip_oracle as so
    from segclip_b200.seg_inference import ViTSegInference, build_text_embedding
    from tools.e2e_report import build_model
    cfg = so.toy_config()
    params = so.init_params(cfg, seed=7)
    model = build_model(cfg, params, "fp32", "torch18_flat").eval()
    g = torch.Generator().manual_seed(9)
    tokens = torch.randint(1, cfg["vocab"] - 2, (5, 2, cfg["context"]), generator=g)
    tokens[:, :, -1] = cfg["vocab"] - 1
    text = build_text_embedding(model, tokens.cuda())
    assert text.shape == (5, cfg["embed_dim"]) and abs(float(text.norm(dim=-1).mean()) - 1) < 1e-4
    img = torch.randn(1, 3, 128, 128, generator=g).cuda()
    whole = ViTSegInference(model, text, True, dict(mode="whole")).inference(img)
    slide1 = ViTSegInference(model, text, True, dict(mode="slide", crop_size=128, stride=128)).inference(img)
    assert torch.allclose(whole, slide1, atol=1e-5)
    seg = ViTSegInference(model, text, True, dict(mode="slide", crop_size=64, stride=32))
    lab = seg.simple_test(img, rescale=True, ori_shape=(150, 140))
    assert lab.shape == (1, 150, 140) and int(lab.max()) < 6 and int(lab.min()) >= 0
