"""Pins oracle/segclip_oracle.py against the UNMODIFIED reference (imported live).

Runs only where /root/reference exists (this container); the GPU box relies on the
committed golden vectors instead (tests/test_oracle_golden.py)."""
import pytest
import torch

from oracle import ref_harness as rh
from oracle import segclip_oracle as so

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("kv_layout", ["torch18_flat", "per_sample"])
@pytest.mark.parametrize("heads_on", [False, True])
def test_toy_loss_and_grads_match_reference(kv_layout, heads_on):
    torch.manual_seed(0)
    cfg = so.toy_config(use_mae=heads_on, use_kl=heads_on)
    B = 20 if kv_layout == "per_sample" else 3   # torch>=2 accepts the reference's call only if B == 8+L
    model = rh.build_reference_model(cfg, kv_layout=kv_layout)
    params = so.init_params(cfg, seed=1)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    batch, noise = so.make_batch(cfg, B, seed=2)
    ref_loss, ref_grads = rh.run_reference(model, batch, noise, heads_on, kv_layout)
    loss, grads, _ = so.loss_and_grads(params, batch, noise, cfg, kv_layout,
                                          frozen=("vis_mae_decoder.decoder_pos_embed",))
    assert abs(float(loss) - float(ref_loss)) <= 2e-5 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    for name, g in ref_grads.items():
        assert name in grads, name
        assert _rel(grads[name], g) < 2e-4, (name, _rel(grads[name], g))
    # parameters the reference leaves without gradient must get none / zero from the port too
    for name, g in grads.items():
        if name not in ref_grads:
            assert float(g.abs().max()) == 0.0, name


@pytest.mark.parametrize("ratio", [0.5, 0.3])
def test_other_mae_mask_ratios_match_reference(ratio):
    """`mae_vis_mask_ratio` other than the recipe's 0.75 (modules/modeling.py:142-145, module_clip_util.py:98): the masked pass
    keeps int(L * (1 - ratio)) tokens."""
    cfg = so.toy_config(use_mae=True, use_kl=True)
    cfg["mae_vis_mask_ratio"] = ratio
    model = rh.build_reference_model(cfg)
    params = so.init_params(cfg, seed=7)
    model.load_state_dict(params, strict=False)
    batch, noise = so.make_batch(cfg, 3, seed=8)
    ref_loss, ref_grads = rh.run_reference(model, batch, noise, True)
    loss, grads, _ = so.loss_and_grads(params, batch, noise, cfg, "torch18_flat", frozen=("vis_mae_decoder.decoder_pos_embed",))
    assert abs(float(loss) - float(ref_loss)) <= 2e-5 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    for name, g in ref_grads.items():
        assert _rel(grads[name], g) < 2e-4, (name, _rel(grads[name], g))


def test_eval_mode_encoders_match_reference():
    """Inference path: eval-mode encode_image / encode_text of the reference (plain softmax assignment,
    modules/module_seg_vit.py:230-231) against the oracle's eval functions."""
    cfg = so.toy_config()
    model = rh.build_reference_model(cfg)
    params = so.init_params(cfg, seed=41)
    model.load_state_dict(params, strict=False)
    model.eval()
    batch, _ = so.make_batch(cfg, 3, seed=42)
    img, ids = batch["image"][:, 0], batch["input_ids"][:, 0]
    with torch.no_grad():
        x, hid, mid = model.clip.encode_image(img, return_hidden=True)
        tx, thid = model.clip.encode_text(ids, return_hidden=True)
        ox, ohid, omid = so.encode_image_eval(img, params, cfg)
        otx, othid = so.encode_text(ids, params, cfg, return_hidden=True)
    model.train()
    for a, b in ((x, ox), (hid, ohid), (mid["hidden"], omid["hidden"]), (tx, otx), (thid, othid),
                 (mid["attns"][0]["soft_attn"], omid["attns"][0]["soft_attn"]),
                 (mid["attns"][0]["hard_attn"], omid["attns"][0]["hard_attn"])):
        assert float((a - b).abs().max()) < 1e-5


@pytest.mark.parametrize("hw", [(128, 128), (64, 256), (256, 64)])
def test_eval_mode_other_resolutions_match_reference(hw):
    """Eval-mode encode_image at input sizes other than the training resolution: the reference interpolates the
    positional table bicubically (modules/module_clip_vtransformer.py:35-53).  SegViT takes its semantic (non-MAE) branch only
    for n or 4 n patch tokens (modules/module_seg_vit.py:423), so the only other sizes the reference supports have 4x the
    tokens: twice the resolution, or a rectangle with the same area.  Toy grid 4 x 4 -> 8 x 8, 4 x 16, 16 x 4."""
    cfg = so.toy_config()
    model = rh.build_reference_model(cfg)
    params = so.init_params(cfg, seed=43)
    model.load_state_dict(params, strict=False)
    model.eval()
    img = torch.randn(2, 3, hw[0], hw[1], generator=torch.Generator().manual_seed(44))
    with torch.no_grad():
        x, hid, mid = model.clip.encode_image(img, return_hidden=True)
        ox, ohid, omid = so.encode_image_eval(img, params, cfg, kv_layout="torch18_flat")
    model.train()
    for a, b in ((x, ox), (hid, ohid), (mid["hidden"], omid["hidden"]),
                 (mid["attns"][0]["soft_attn"], omid["attns"][0]["soft_attn"])):
        assert a.shape == b.shape and float((a - b).abs().max()) < 1e-5


def test_reference_copy_for_the_gpu_box_is_untouched(tmp_path):
    """oracle/make_ref.py (run by __graft_entry__.build()) places an UNMODIFIED copy of the reference's Python modules under
    oracle/_ref for bench.py's reference arm on the GPU box: byte-identical files, git-ignored."""
    import filecmp
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "oracle", "make_ref.py")], check=True, capture_output=True)
    ref = "/root/reference"
    for f in [os.path.join("modules", x) for x in os.listdir(os.path.join(ref, "modules")) if x.endswith(".py")] + ["util.py"]:
        assert filecmp.cmp(os.path.join(ref, f), os.path.join(root, "oracle", "_ref", f), shallow=False), f
    ignored = subprocess.run(["git", "check-ignore", "oracle/_ref/util.py"], cwd=root, capture_output=True, text=True)
    assert ignored.returncode == 0, "oracle/_ref must stay out of the repository history"
