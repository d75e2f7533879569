"""Per-kernel parity of the non-GEMM entry points (through the C ABI) against the CPU oracle's own functions and their
autograd gradients: LayerNorm (+row remap), assignment / aggregation incl. empty centres, ReconstructLayer, superpixel KL,
InfoNCE with rank-offset labels and gathered LSEs, MAE masking / un-shuffle / loss, pooling."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import segclip_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("D", [64, 384, 768, 1024])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(D, dtype):
    from segclip_b200 import ops
    torch.manual_seed(D)
    rows = 203                                    # ragged: not a multiple of the 8 rows per CTA
    x = torch.randn(rows, D) * 2 + 0.5
    g, b = torch.randn(D), torch.randn(D)
    dy = torch.randn(rows, D).to(dtype)
    base = torch.randn(rows, D)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = F.layer_norm(xr, (D,), gr, br, 1e-5)
    y_ref.backward(dy.float())
    xd, gd, bd = x.to(DEV), g.to(DEV), b.to(DEV)
    y = torch.empty(rows, D, device=DEV, dtype=dtype)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.layernorm_op(xd, gd, bd, y, 1e-5, mean, rstd)()
    dx = base.to(DEV).clone()
    dxc = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.layernorm_bwd_op(dy.to(DEV), xd, mean, rstd, gd, dx, True, dxc, dg, db)()
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert _rel(y, y_ref.detach()) < tol
    assert _rel(dx, base + xr.grad) < 1e-4
    assert _rel(dxc, base + xr.grad) < 1e-2
    assert _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("D", [512, 768, 1024, 384, 132])
def test_layernorm_bwd_bf16_gradient_stream_with_colsum(D):
    """The production configuration of the big streams: dy bf16, x fp32, the gradient stream dx accumulated IN PLACE in
    bf16 (no fp32 copy), dgamma / dbeta and the column sums of the final dx (bias gradient of the residual branch's Linear)
    from register accumulators.  D = 132 is not a multiple of 8 -> the shared-memory fallback kernel."""
    from segclip_b200 import ops
    torch.manual_seed(D)
    rows = 1003
    x = (torch.randn(rows, D) * 2 + 0.5).to(DEV)
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    dy = torch.randn(rows, D, device=DEV).bfloat16()
    base = torch.randn(rows, D, device=DEV).bfloat16()
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gr, br, 1e-5).backward(dy.float())
    want = base.float() + xr.grad
    y = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.layernorm_op(x, g, b, y, 1e-5, mean, rstd)()
    dx = base.clone()
    dg, db, cs = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.layernorm_bwd_op(dy, x, mean, rstd, g, dx, True, None, dg, db, dx_colsum=cs)()
    torch.cuda.synchronize()
    assert _rel(dx, want) < 1e-2
    assert _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4
    assert _rel(cs, want.sum(0)) < 1e-4                      # summed before the bf16 rounding of the store
    # not accumulating, no parameter gradients (cross-attention K/V side): plain overwrite
    dx2 = torch.full_like(dx, float("nan"))
    ops.layernorm_bwd_op(dy, x, mean, rstd, g, dx2, False)()
    assert _rel(dx2, xr.grad) < 1e-2


def test_layernorm_row_remap_into_concat_buffer():
    """kv_ = cat([q_feat, x], 1) (module_seg_vit.py:294) is never materialised: both LayerNorms write into it."""
    from segclip_b200 import ops
    torch.manual_seed(0)
    B, G, Lx, D = 3, 8, 5, 128
    q, x = torch.randn(B * G, D, device=DEV), torch.randn(B * Lx, D, device=DEV)
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    out = torch.full((B * (G + Lx), D), float("nan"), device=DEV)
    ops.layernorm_op(q, g, b, out, remap=(G, G + Lx, 0))()
    ops.layernorm_op(x, g, b, out, remap=(Lx, G + Lx, G))()
    ref = F.layer_norm(torch.cat([q.view(B, G, D), x.view(B, Lx, D)], 1), (D,), g, b).view(-1, D)
    assert _rel(out, ref) < 1e-5


@pytest.mark.parametrize("empty_center", [False, True])
def test_assignment_aggregation_fwd_bwd(empty_center):
    """Hard Gumbel assignment + weighted mean + clamp normaliser and their straight-through backward
    (module_seg_vit.py:221-242,304-310) against autograd of the oracle's formulation."""
    from segclip_b200 import ops
    torch.manual_seed(3)
    B, G, Lx, D = 2, 8, 19, 64
    qf, k, v = torch.randn(B, G, D), torch.randn(B, Lx, D), torch.randn(B, Lx, D)
    if empty_center:                              # centre 5 never wins -> count 0 -> clamp_min(., 1) path
        k[..., 0] = k[..., 0].abs() + 5.0
        qf[:, 5] = 0.0
        qf[:, 5, 0] = -100.0
    u = torch.rand(B, G, Lx)
    extra = torch.randn(B, G, Lx) * 0.1
    d_out = torch.randn(B, G, D)
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (qf, k, v))
    attn = torch.einsum("bgc,blc->bgl", qr, kr)
    y = torch.softmax((attn + so.gumbel_from_uniform(u)) / 0.9, dim=1)
    idx_ref = y.argmax(1)
    hard = torch.zeros_like(y).scatter_(1, idx_ref.unsqueeze(1), 1.0) - y.detach() + y
    out = torch.einsum("bgl,blc->bgc", hard, vr) / torch.clamp_min(hard.sum(-1, keepdim=True), 1.0)
    total = qr + out
    (total * d_out).sum().backward(retain_graph=True)
    (hard * extra).sum().backward()
    # native
    qd, kd, vd, ud = qf.to(DEV), k.to(DEV), v.to(DEV), u.to(DEV)
    y_soft, soft = torch.empty(B, G, Lx, device=DEV), torch.empty(B, G, Lx, device=DEV)
    idx, count = torch.empty(B, Lx, device=DEV, dtype=torch.int32), torch.zeros(B, G, device=DEV)
    ops.assign_fwd_op(qd.view(-1, D), kd.view(-1, D), ud, y_soft, idx, count, B, Lx, D, 0.9, None, soft)()
    agg, ssum = torch.empty(B * G, D, device=DEV), torch.empty(B * G, D, device=DEV)
    ops.aggregate_fwd_op(vd.view(-1, D), idx, count, qd.view(-1, D), agg, ssum, B, Lx, D)()
    d_logits, d_v, d_k = torch.empty(B, G, Lx, device=DEV), torch.empty(B * Lx, D, device=DEV), torch.empty(B * Lx, D, device=DEV)
    d_qf = torch.empty(B * G, D, device=DEV)
    dsum = d_out.to(DEV).view(-1, D).contiguous()
    ops.assign_bwd_op(dsum, agg, vd.view(-1, D), idx, count, y_soft, extra.to(DEV), qd.view(-1, D), kd.view(-1, D), d_logits, d_v,
                      d_k, dsum, d_qf, B, Lx, D, 0.9)()
    torch.cuda.synchronize()
    assert torch.equal(idx.cpu().long(), idx_ref)
    if empty_center:
        assert float(count[:, 5].sum()) == 0.0
    assert _rel(soft, torch.softmax(attn, 1).detach()) < 1e-5
    assert _rel(ssum.view(B, G, D), total.detach()) < 1e-5
    assert _rel(d_qf.view(B, G, D), qr.grad) < 1e-4
    assert _rel(d_k.view(B, Lx, D), kr.grad) < 1e-4
    assert _rel(d_v.view(B, Lx, D), vr.grad) < 1e-4


@pytest.mark.parametrize("B,Lx,D,vdt", [(3, 196, 768, torch.bfloat16), (2, 48, 768, torch.float32), (5, 19, 128, torch.bfloat16),
                                        (2, 256, 1024, torch.bfloat16), (1, 784, 768, torch.float32)])
def test_fused_assign_aggregate_kernel(B, Lx, D, vdt):
    """sc_assign_aggregate_fwd (assignment softmax + arg-max + weighted mean in ONE kernel, one CTA per sample) against the
    two-kernel path, and the one-kernel backward behind sc_assign_bwd against autograd of the oracle's formulation."""
    from segclip_b200 import ops
    torch.manual_seed(B * Lx + D)
    G = 8
    qf, k = torch.randn(B, G, D) / D ** 0.25, torch.randn(B, Lx, D) / D ** 0.25
    v = torch.randn(B, Lx, D).to(vdt).float()
    u, extra, d_out = torch.rand(B, G, Lx), torch.randn(B, G, Lx) * 0.1, torch.randn(B, G, D)
    qd, kd, vd, ud = qf.to(DEV).view(-1, D), k.to(DEV).view(-1, D), v.to(DEV).to(vdt).view(-1, D), u.to(DEV)

    def outs():
        return dict(y=torch.empty(B, G, Lx, device=DEV), soft=torch.empty(B, G, Lx, device=DEV),
                    idx=torch.empty(B, Lx, device=DEV, dtype=torch.int32), count=torch.zeros(B, G, device=DEV),
                    agg=torch.empty(B * G, D, device=DEV), ssum=torch.empty(B * G, D, device=DEV))
    a, f = outs(), outs()
    ops.assign_fwd_op(qd, kd, ud, a["y"], a["idx"], a["count"], B, Lx, D, 0.9, None, a["soft"])()
    ops.aggregate_fwd_op(vd, a["idx"], a["count"], qd, a["agg"], a["ssum"], B, Lx, D)()
    f["count"].fill_(7.0)                                    # written, not accumulated
    ops.assign_aggregate_fwd_op(qd, kd, ud, f["y"], f["idx"], f["count"], vd, f["agg"], f["ssum"], B, Lx, D, 0.9, None, f["soft"])()
    torch.cuda.synchronize()
    flips = float((a["idx"] != f["idx"]).float().mean())
    assert flips == 0.0, flips
    assert torch.equal(a["count"], f["count"])
    for key in ("y", "soft", "agg", "ssum"):
        assert _rel(f[key], a[key]) < 1e-5, key
    # inference variant (no Gumbel noise) and teacher forcing
    forced = torch.randint(0, G, (B, Lx), device=DEV, dtype=torch.int32)
    e = outs()
    ops.assign_aggregate_fwd_op(qd, kd, None, e["y"], e["idx"], e["count"], vd, e["agg"], e["ssum"], B, Lx, D, 0.9, forced, e["soft"])()
    torch.cuda.synchronize()
    assert torch.equal(e["idx"], forced)
    hard = torch.zeros(B, G, Lx, device=DEV).scatter_(1, forced.long().unsqueeze(1), 1.0)
    want = torch.einsum("bgl,blc->bgc", hard, vd.float().view(B, Lx, D)) / hard.sum(-1, keepdim=True).clamp_min(1.0)
    assert _rel(e["agg"].view(B, G, D), want) < 1e-5
    # backward (fused kernel behind sc_assign_bwd) vs autograd
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (qf, k, v))
    attn = torch.einsum("bgc,blc->bgl", qr, kr)
    y = torch.softmax((attn + so.gumbel_from_uniform(u)) / 0.9, dim=1)
    idx_ref = f["idx"].cpu().long()                          # the kernel's own arg-max (equal to the reference's up to exact ties)
    hard = torch.zeros_like(y).scatter_(1, idx_ref.unsqueeze(1), 1.0) - y.detach() + y
    out = torch.einsum("bgl,blc->bgc", hard, vr) / torch.clamp_min(hard.sum(-1, keepdim=True), 1.0)
    ((qr + out) * d_out).sum().backward(retain_graph=True)
    (hard * extra).sum().backward()
    d_logits, d_v = torch.empty(B, G, Lx, device=DEV), torch.empty(B * Lx, D, device=DEV, dtype=vdt)
    d_k, d_qf = torch.empty(B * Lx, D, device=DEV), torch.empty(B * G, D, device=DEV)
    dsum = d_out.to(DEV).view(-1, D).contiguous()
    ops.assign_bwd_op(dsum, f["agg"], vd, f["idx"], f["count"], f["y"], extra.to(DEV), qd, kd, d_logits, d_v, d_k, dsum, d_qf,
                      B, Lx, D, 0.9)()
    torch.cuda.synchronize()
    assert _rel(d_qf.view(B, G, D), qr.grad) < 1e-4
    assert _rel(d_k.view(B, Lx, D), kr.grad) < 1e-4
    assert _rel(d_v.view(B, Lx, D), vr.grad) < (1e-2 if vdt == torch.bfloat16 else 1e-4)


def test_reconstruct_layer_fwd_bwd():
    from segclip_b200 import ops
    torch.manual_seed(4)
    B, G, M, D = 3, 8, 11, 64
    sx, W, bias = torch.randn(B, G, D), torch.randn(G, G) * 0.3, torch.randn(G) * 0.1
    idx = torch.randint(0, G, (B, M))
    d_out = torch.randn(B, M, D)
    sxr, Wr, br = (t.clone().requires_grad_(True) for t in (sx, W, bias))
    hard = torch.zeros(B, G, M).scatter_(1, idx.unsqueeze(1), 1.0).requires_grad_(True)
    p = {"r.rec_proj_a.a_fc.weight": Wr, "r.rec_proj_a.a_fc.bias": br}
    out_ref = so.reconstruct_layer(sxr, hard, p, "r.")
    (out_ref * d_out).sum().backward()
    pre, out = torch.empty(B * M, D, device=DEV), torch.empty(B * M, D, device=DEV)
    idd = idx.to(DEV).int()
    ops.reconstruct_fwd_op(sx.to(DEV).view(-1, D), idd, W.to(DEV), bias.to(DEV), pre, out, B, M, D)()
    d_sx, d_hard = torch.empty(B * G, D, device=DEV), torch.empty(B, G, M, device=DEV)
    dW, db = torch.zeros(G, G, device=DEV), torch.zeros(G, device=DEV)
    ops.reconstruct_bwd_op(d_out.to(DEV).view(-1, D).contiguous(), pre, sx.to(DEV).view(-1, D), idd, W.to(DEV), bias.to(DEV), d_sx,
                           d_hard, dW, db, B, M, D)()
    torch.cuda.synchronize()
    assert _rel(out.view(B, M, D), out_ref.detach()) < 1e-5
    assert _rel(d_sx.view(B, G, D), sxr.grad) < 1e-4
    assert _rel(d_hard, hard.grad) < 1e-4
    assert _rel(dW, Wr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4


def test_superpixel_kl_loss_and_gradient():
    """modules/modeling.py:212-224, incl. labels that are arbitrary int64 values and singleton segments."""
    from segclip_b200 import ops
    torch.manual_seed(5)
    B, G, Lx = 3, 8, 25
    idx = torch.randint(0, G, (B, Lx))
    seg = torch.randint(0, 4, (B, Lx)) * 1000003 - 7          # arbitrary labels
    seg[0, 3] = 999999999                                      # a singleton segment
    hard = torch.zeros(B, G, Lx).scatter_(1, idx.unsqueeze(1), 1.0).requires_grad_(True)
    ref = so.superpixel_kl(hard, seg.view(B, 5, 5))
    ref.backward()
    loss, d_hard = torch.zeros(1, device=DEV), torch.empty(B, G, Lx, device=DEV)
    ops.superpixel_kl_op(idx.to(DEV).int(), seg.to(DEV), loss, d_hard, B, Lx)()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 1e-6 + 1e-5 * abs(float(ref))
    assert _rel(d_hard, hard.grad) < 1e-4


@pytest.mark.parametrize("world,rank", [(1, 0), (4, 2)])
def test_infonce_rank_offset_labels_and_lse_backward(world, rank):
    """One rank's view of the W-rank InfoNCE: own-row cross-entropy forward, and the reduction-free backward that uses the
    gathered log-sum-exps of the transposed direction -- checked against autograd of the SUM of all ranks'
    losses w.r.t. this rank's embeddings (what diffdist's reduce-scatter delivers)."""
    from segclip_b200 import ops
    torch.manual_seed(6)
    B, E = 5, 32
    N = B * world
    t_all = F.normalize(torch.randn(N, E), dim=-1).requires_grad_(True)
    v_all = F.normalize(torch.randn(N, E), dim=-1).requires_grad_(True)
    p = torch.tensor(math.log(20.0), requires_grad=True)
    losses = [so.contrastive_loss(t_all[r * B:(r + 1) * B], v_all[r * B:(r + 1) * B], t_all, v_all, p, r) for r in range(world)]
    gt, gv = torch.autograd.grad(sum(losses), (t_all, v_all), retain_graph=True)
    gp, = torch.autograd.grad(losses[rank], p)
    lo = rank * B
    s = float(p.exp())
    lse_t2v = torch.logsumexp(s * t_all.detach() @ v_all.detach().t(), 1)      # global rows
    lse_v2t = torch.logsumexp(s * v_all.detach() @ t_all.detach().t(), 1)
    td, vd = t_all.detach().to(DEV), v_all.detach().to(DEV)
    raw_t2v, raw_v2t = (td[lo:lo + B] @ vd.t()).contiguous(), (vd[lo:lo + B] @ td.t()).contiguous()
    pd = p.detach().to(DEV).view(1)
    loss, lse_a, lse_b = torch.zeros(1, device=DEV), torch.empty(B, device=DEV), torch.empty(B, device=DEV)
    ops.ce_lse_op(raw_t2v, lo, pd, lse_a, loss)()
    ops.ce_lse_op(raw_v2t, lo, pd, lse_b, loss)()
    dscale = torch.zeros(1, device=DEV)
    ops.ce_grad_op(raw_t2v, lo, pd, lse_a, lse_v2t.to(DEV), dscale)()
    ops.ce_grad_op(raw_v2t, lo, pd, lse_b, lse_t2v.to(DEV), dscale)()
    d_t, d_v = raw_t2v @ vd, raw_v2t @ td
    torch.cuda.synchronize()
    assert abs(float(loss) - float(losses[rank])) < 1e-5 * abs(float(losses[rank]))
    assert _rel(lse_a, lse_t2v[lo:lo + B]) < 1e-5
    assert _rel(d_t, gt[lo:lo + B]) < 1e-4 and _rel(d_v, gv[lo:lo + B]) < 1e-4
    assert abs(float(dscale) - float(gp)) < 1e-4 * abs(float(gp)) + 1e-7


def test_logit_scale_clamp_gradient_is_zero_above_100():
    from segclip_b200 import ops
    B, E = 4, 16
    t = F.normalize(torch.randn(B, E), dim=-1).to(DEV)
    raw = (t @ t.t()).contiguous()
    p = torch.tensor([math.log(150.0)], device=DEV)           # exp(p) > 100 -> clamp (modeling.py:350) -> no gradient
    loss, lse, d = torch.zeros(1, device=DEV), torch.empty(B, device=DEV), torch.zeros(1, device=DEV)
    ops.ce_lse_op(raw, 0, p, lse, loss)()
    ops.ce_grad_op(raw, 0, p, lse, lse, d)()
    torch.cuda.synchronize()
    ref = torch.logsumexp(100.0 * (t @ t.t()), 1)
    assert _rel(lse, ref) < 1e-5 and float(d) == 0.0


def test_mae_masking_unshuffle_and_loss():
    from segclip_b200 import ops
    torch.manual_seed(7)
    cfg = so.toy_config(use_mae=True)
    B, grid, patch = 3, cfg["grid"], cfg["patch"]
    L1, keep, dd = grid * grid + 1, int((grid * grid + 1) * 0.25), 64
    u = torch.rand(B, L1)
    x = torch.randn(B, L1, 8)
    _, mask_ref, restore_ref, keep_ref = so.random_masking_keep_cls(x, u)
    ids_restore, ids_keep = torch.empty(B, L1, device=DEV, dtype=torch.int32), torch.empty(B, keep, device=DEV, dtype=torch.int32)
    mask, pidx = torch.empty(B, L1, device=DEV), torch.empty(B * (keep - 1), device=DEV, dtype=torch.int32)
    ops.mae_mask_op(u.to(DEV), ids_restore, ids_keep, mask, pidx, B, L1, keep)()
    torch.cuda.synchronize()
    assert torch.equal(ids_restore.cpu().long(), restore_ref) and torch.equal(ids_keep.cpu().long(), keep_ref)
    assert torch.equal(mask.cpu(), mask_ref)
    assert torch.equal(pidx.cpu().long().view(B, keep - 1), keep_ref[:, 1:] - 1)
    # un-shuffle + decoder positional table, and its backward
    emb = torch.randn(B, keep, dd, requires_grad=True)
    mtok, pos = torch.randn(1, 1, dd, requires_grad=True), torch.randn(1, L1, dd)
    xx = torch.cat([emb, mtok.expand(B, L1 - keep, -1)], 1)
    xx = torch.gather(xx, 1, restore_ref.unsqueeze(-1).expand(-1, -1, dd)) + pos
    dxx = torch.randn(B, L1, dd)
    (xx * dxx).sum().backward()
    out = torch.empty(B * L1, dd, device=DEV)
    ops.mae_unshuffle_op(emb.detach().to(DEV).view(-1, dd), mtok.detach().to(DEV).view(-1), ids_restore, pos.to(DEV).view(-1, dd), out,
                         B, L1, keep, dd)()
    d_emb, d_tok = torch.empty(B * keep, dd, device=DEV), torch.zeros(dd, device=DEV)
    ops.mae_unshuffle_bwd_op(dxx.to(DEV).view(-1, dd).contiguous(), ids_restore, d_emb, d_tok, B, L1, keep, dd)()
    torch.cuda.synchronize()
    assert _rel(out.view(B, L1, dd), xx.detach()) < 1e-6
    assert _rel(d_emb.view(B, keep, dd), emb.grad) < 1e-6 and _rel(d_tok, mtok.grad.view(-1)) < 1e-5
    # masked-patch MSE against patchify(image)
    P = 3 * patch * patch
    img = torch.randn(B, 3, grid * patch, grid * patch)
    pred = torch.randn(B, L1, P, requires_grad=True)
    per_patch = ((pred[:, 1:] - so.patchify(img, patch)) ** 2).mean(-1)
    ref = (per_patch * mask_ref[:, 1:]).sum() / mask_ref[:, 1:].sum()
    ref.backward()
    loss, dpred = torch.zeros(1, device=DEV), torch.empty(B * L1, P, device=DEV)
    ops.mae_loss_op(pred.detach().to(DEV).view(-1, P), img.to(DEV), mask, loss, dpred, B, L1, keep, grid, patch)()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert _rel(dpred.view(B, L1, P), pred.grad) < 1e-5


def test_pooling_kernels():
    from segclip_b200 import ops
    torch.manual_seed(8)
    B, G, D, n = 4, 8, 96, 5
    x = torch.randn(B, G, D, requires_grad=True)
    m, arg_ref = x.max(1)
    dm = torch.randn(B, D)
    (m * dm).sum().backward()
    out, arg, dx = torch.empty(B, D, device=DEV), torch.empty(B, D, device=DEV, dtype=torch.int32), torch.empty(B * G, D, device=DEV)
    ops.pool_max_op(x.detach().to(DEV).view(-1, D), out, arg, B, G, D)()
    ops.pool_max_bwd_op(dm.to(DEV), arg, dx, B, G, D)()
    y = torch.randn(B, n, D, requires_grad=True)
    cat = torch.cat([y.mean(1, keepdim=True), y], 1)
    dcat = torch.randn(B, n + 1, D)
    (cat * dcat).sum().backward()
    o2, dy = torch.empty(B * (n + 1), D, device=DEV), torch.empty(B * n, D, device=DEV)
    ops.mean_cat_op(y.detach().to(DEV).view(-1, D), o2, B, n, D)()
    ops.mean_cat_bwd_op(dcat.to(DEV).view(-1, D).contiguous(), dy, B, n, D)()
    torch.cuda.synchronize()
    assert torch.equal(arg.cpu().long(), arg_ref) and _rel(out, m.detach()) == 0.0
    assert _rel(dx.view(B, G, D), x.grad) < 1e-6
    assert _rel(o2.view(B, n + 1, D), cat.detach()) < 1e-6 and _rel(dy.view(B, n, D), y.grad) < 1e-6
