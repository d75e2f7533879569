"""sc_gemm (tcgen05 bf16 kernel and fp32 FMA kernel) against a plain fp32 torch reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, ta, tb, bias, rowbias, ridx, rmod, act, residual, alpha):
    a = A.float().t() if ta else A.float()
    b = B.float().t() if tb else B.float()
    v = alpha * (a.double() @ b.double().t()).float()
    if bias is not None:
        v = v + bias
    if rowbias is not None:
        idx = ridx.long() if ridx is not None else torch.arange(v.shape[0], device=v.device) % rmod
        v = v + rowbias[idx]
    pre = v
    if act == 1:
        v = v * torch.sigmoid(1.702 * v)
    elif act == 2:
        v = torch.nn.functional.gelu(v)
    if residual is not None:
        v = v + residual.float()
    return v, pre


def _run(M, N, K, dtype, ta=False, tb=False, bias=False, rowbias=0, ridx=False, act=0, residual=False,
         c_dtype=torch.float32, c2=None, accumulate=False, split_k=0, alpha=1.0, force_simt=False, seed=0, res_dtype=torch.float32):
    from segclip_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    A = torch.randn((K, M) if ta else (M, K), device=dev, generator=g).to(dtype)
    B = torch.randn((K, N) if tb else (N, K), device=dev, generator=g).to(dtype)
    bias_t = torch.randn(N, device=dev, generator=g) if bias else None
    rb = torch.randn(rowbias, N, device=dev, generator=g) if rowbias else None
    ridx_t = torch.randint(0, rowbias, (M,), device=dev, generator=g, dtype=torch.int32) if (ridx and rowbias) else None
    res = torch.randn(M, N, device=dev, generator=g).to(res_dtype) if residual else None
    C0 = torch.randn(M, N, device=dev, generator=g) if accumulate else None
    C = C0.clone() if accumulate else torch.full((M, N), float("nan"), device=dev, dtype=c_dtype)
    C2 = torch.full((M, N), float("nan"), device=dev, dtype=c2) if c2 is not None else None
    ops.gemm(A, B, C, trans_a=ta, trans_b=tb, bias=bias_t, rowbias=rb, rowbias_idx=ridx_t, act=act, residual=res,
             C2=C2, accumulate=accumulate, split_k=split_k, alpha=alpha, force_simt=force_simt)
    torch.cuda.synchronize()
    want, pre = _ref(A, B, ta, tb, bias_t, rb, ridx_t, rowbias, act, res, alpha)
    if accumulate:
        want = want + C0
    scale = float(want.abs().max()) + 1e-6
    tol = 2e-5 if (dtype == torch.float32 and c_dtype == torch.float32) else (1e-2 if c_dtype == torch.bfloat16 else 2e-4)
    err = float((C.float() - want).abs().max()) / scale
    assert err < tol, (err, M, N, K, ta, tb)
    if C2 is not None:
        err2 = float((C2.float() - pre).abs().max()) / (float(pre.abs().max()) + 1e-6)
        assert err2 < (1e-2 if c2 == torch.bfloat16 else 2e-4), err2


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 128), (128, 128, 64), (392, 768, 768), (300, 384, 200),
                                   (50, 64, 72), (1000, 2304, 768), (16, 512, 512)])
def test_tc_nt_plain(M, N, K):
    _run(M, N, K, torch.bfloat16)


@pytest.mark.parametrize("ta,tb", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 192), (768, 768, 392), (3072, 768, 1000), (200, 328, 136)])
def test_tc_transposed_operands(M, N, K, ta, tb):
    _run(M, N, K, torch.bfloat16, ta=ta, tb=tb)


def test_tc_epilogues():
    _run(392, 3072, 768, torch.bfloat16, bias=True, act=1, c_dtype=torch.bfloat16, c2=torch.bfloat16)
    _run(392, 768, 3072, torch.bfloat16, bias=True, residual=True)
    _run(392, 768, 768, torch.bfloat16, rowbias=196)
    _run(96, 768, 768, torch.bfloat16, rowbias=196, ridx=True, c_dtype=torch.bfloat16)
    _run(16, 3072, 768, torch.bfloat16, bias=True, act=2, c_dtype=torch.bfloat16, c2=torch.bfloat16)
    _run(256, 512, 512, torch.bfloat16, alpha=0.5, accumulate=True)


def test_tc_split_k():
    _run(768, 768, 8192, torch.bfloat16, ta=True, tb=True, accumulate=True, split_k=8)
    _run(2304, 768, 5000, torch.bfloat16, ta=True, tb=True, accumulate=True, split_k=-1)
    _run(768, 3072, 1568, torch.bfloat16, ta=True, tb=True, accumulate=True, split_k=-1)


def test_tc_many_tiles_persistent():
    # more tiles than SMs: exercises the smem ring / TMEM double-buffer phase logic
    _run(8192, 2304, 768, torch.bfloat16, bias=True, c_dtype=torch.bfloat16)
    _run(6272, 3072, 768, torch.bfloat16, bias=True, act=1, c_dtype=torch.bfloat16)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, True)])
def test_simt_fp32(ta, tb):
    _run(200, 136, 77, torch.float32, ta=ta, tb=tb, bias=True, act=1, residual=True, c2=torch.float32)
    _run(64, 8, 8, torch.float32, ta=ta, tb=tb, bias=True)


def test_tc_matches_simt_on_bf16_inputs():
    _run(300, 384, 200, torch.bfloat16, force_simt=True, bias=True)


def test_tc_fused_activation_backward():
    """dgrad with the fused act'(pre) multiplier (QuickGELU / erf-GELU backward folded into the epilogue)."""
    from segclip_b200 import ops
    torch.manual_seed(0)
    for act in (1, 2):
        M, N, K = 392, 3072, 768
        dy = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
        pre = torch.randn(M, N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(dy, W, out, trans_b=True, mul_aux=pre, mul_aux_act=act)
        x = pre.float().requires_grad_(True)
        y = x * torch.sigmoid(1.702 * x) if act == 1 else torch.nn.functional.gelu(x)
        y.backward(dy.float() @ W.float())
        err = float((out.float() - x.grad).abs().max() / x.grad.abs().max())
        assert err < 1e-2, err


def test_colsum_vectorised_and_scalar():
    from segclip_b200 import ops
    torch.manual_seed(0)
    for rows, cols, dt in [(5000, 3072, torch.bfloat16), (777, 768, torch.float32), (100, 36, torch.float32)]:
        x = torch.randn(rows, cols, device="cuda").to(dt)
        out = torch.zeros(cols, device="cuda")
        ops.colsum_op(x, out)()
        ref = x.float().sum(0)
        assert float((out - ref).abs().max()) < 1e-3 * (1 + float(ref.abs().max()))


# ---- 2-CTA kernel (M >= 512, N >= 256): TMA-store epilogues, ragged edges clipped by the tensor map -----------------
@pytest.mark.parametrize("M,N,K", [(1000, 3072, 768), (2048, 512, 512), (777, 1000, 256), (4100, 2304, 768)])
def test_tc2_tma_store_epilogues(M, N, K):
    _run(M, N, K, torch.bfloat16, c_dtype=torch.bfloat16)                                              # plain
    _run(M, N, K, torch.bfloat16, bias=True, c_dtype=torch.bfloat16)                                   # + bias
    _run(M, N, K, torch.bfloat16, bias=True, act=1, c_dtype=torch.bfloat16, c2=torch.bfloat16)         # QuickGELU, two outputs
    _run(M, N, K, torch.bfloat16, bias=True, act=2, c_dtype=torch.bfloat16, c2=torch.bfloat16)         # erf-GELU, two outputs
    _run(M, N, K, torch.bfloat16, tb=True, c_dtype=torch.bfloat16)                                     # dgrad layout
    _run(M, N, K, torch.bfloat16, bias=True, residual=True)                                            # fp32 residual path


def test_tc2_outputs_do_not_spill_past_the_edges():
    """TMA boxes that overhang M or N must not write outside C (checked with a guard band around a strided C)."""
    from segclip_b200 import ops
    torch.manual_seed(0)
    M, N, K, pad = 1000, 1000, 256, 8
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    big = torch.full((M + 40, N + pad), 7.0, device="cuda", dtype=torch.bfloat16)
    C = big[:M, :N]
    ops.gemm(A, B, C, bias=torch.zeros(N, device="cuda"))
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert float((C.float() - ref).abs().max() / ref.abs().max()) < 1e-2
    assert bool((big[M:] == 7.0).all()) and bool((big[:, N:] == 7.0).all())


def test_tc2_fused_activation_backward_with_colsum():
    """The activation-gradient dgrad of the 2-CTA kernel incl. the fused column sums (c_fc bias gradient)."""
    from segclip_b200 import ops
    torch.manual_seed(0)
    for act in (1, 2):
        M, N, K = 1576, 3072, 768
        dy = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
        pre = torch.randn(M, N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        cs = torch.zeros(N, device="cuda")
        ops.gemm(dy, W, out, trans_b=True, mul_aux=pre, mul_aux_act=act, colsum_out=cs)
        torch.cuda.synchronize()
        x = pre.float().requires_grad_(True)
        y = x * torch.sigmoid(1.702 * x) if act == 1 else torch.nn.functional.gelu(x)
        y.backward(dy.float() @ W.float())
        assert float((out.float() - x.grad).abs().max() / x.grad.abs().max()) < 1e-2
        ref = out.float().sum(0)
        assert float((cs - ref).abs().max()) < 2e-3 * (1 + float(ref.abs().max()))


@pytest.mark.parametrize("M,N,K", [(392, 768, 768), (1000, 768, 768), (4100, 768, 1536)])
def test_tc_plain_fp32_output_and_accumulate(M, N, K):
    """Specialised fp32-output epilogues of both tcgen05 kernels: plain store and non-atomic C += (alpha = 1)."""
    _run(M, N, K, torch.bfloat16)                                   # NN, fp32 out
    _run(M, N, K, torch.bfloat16, tb=True)                          # NT, fp32 out
    _run(M, N, K, torch.bfloat16, tb=True, accumulate=True)         # NT, C += (old C prefetched)
    _run(M, N, K, torch.bfloat16, accumulate=True)                  # NN accumulate: generic path


@pytest.mark.parametrize("M,N,K", [(1576, 768, 768), (4100, 768, 3072), (600, 512, 2048), (392, 768, 768), (154, 512, 512)])
def test_tc_bf16_residual_stream_epilogue(M, N, K):
    """out_proj / c_proj forward with the bf16 residual stream: C(bf16) = A B^T + bias + residual(bf16).  M >= 512 runs the
    2-CTA kernel (residual tile TMA-loaded into the staging box its result is TMA-stored from, ragged last row tile
    included), smaller M the 1-CTA kernel's generic epilogue."""
    from segclip_b200 import _lib
    k0 = _lib.kernel_launches()
    _run(M, N, K, torch.bfloat16, bias=True, residual=True, c_dtype=torch.bfloat16, res_dtype=torch.bfloat16)
    k1 = _lib.kernel_launches()
    assert (k1["gemm_tc2"] - k0["gemm_tc2"] == 1) == (M >= 512), (k0, k1)


def test_tc2_residual_epilogue_in_place_and_guard_band():
    """The residual may alias the output (x += f(x) in place is not used by the engine, but the tile is read before it is
    written); rows past M must stay untouched."""
    from segclip_b200 import ops
    torch.manual_seed(1)
    M, N, K = 1000, 768, 256
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    big = torch.full((M + 40, N), 7.0, device="cuda", dtype=torch.bfloat16)
    C = big[:M]
    ops.gemm(A, B, C, bias=bias, residual=res)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + bias + res.float()
    assert float((C.float() - ref).abs().max() / ref.abs().max()) < 1e-2
    assert bool((big[M:] == 7.0).all())


@pytest.mark.parametrize("M,N,K", [(1576, 768, 768), (4100, 512, 2048), (1000, 768, 3072), (50176, 768, 768)])
def test_tc2_fp32_residual_stream_epilogue(M, N, K):
    """out_proj / c_proj forward of the 2-CTA kernel: C(fp32) = A B^T + bias + residual(fp32), residual in and result out
    through the TMA (two 4 KB fp32 boxes per epilogue warp, reloaded inside the tile), ragged last row tile included."""
    from segclip_b200 import _lib
    k0 = _lib.kernel_launches()
    _run(M, N, K, torch.bfloat16, bias=True, residual=True)
    assert _lib.kernel_launches()["gemm_tc2"] - k0["gemm_tc2"] == 1


def test_tc2_fp32_residual_guard_band():
    from segclip_b200 import ops
    torch.manual_seed(2)
    M, N, K = 1000, 768, 256
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    big = torch.full((M + 40, N), 7.0, device="cuda")
    C = big[:M]
    ops.gemm(A, B, C, bias=bias, residual=res)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + bias + res
    assert float((C - ref).abs().max() / ref.abs().max()) < 2e-4
    assert bool((big[M:] == 7.0).all())


@pytest.mark.parametrize("B,L", [(8, 197), (2, 196), (6, 77), (40, 196)])
def test_dgrad_with_fused_attention_delta(B, L):
    """out_proj dgrad dO = dY W with the per-head row dots delta[b,h,l] = sum_d dO[b,l,h,d] O[b,l,h,d] fused into the
    epilogue (2-CTA kernel: O tile TMA-loaded into the staging box; M < 512: stand-alone pass after the 1-CTA GEMM)."""
    from segclip_b200 import ops
    torch.manual_seed(B)
    M, N, K = B * L, 768, 768
    H = N // 64
    dy = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    O = torch.randn(M, N, device="cuda").bfloat16()
    dO = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    delta = torch.full((B, H, L), float("nan"), device="cuda")
    ops.gemm(dy, W, dO, trans_b=True, dot_aux=O, dot_out=delta, dot_L=L)
    torch.cuda.synchronize()
    ref = dy.float() @ W.float()
    assert float((dO.float() - ref).abs().max() / ref.abs().max()) < 1e-2
    want = (ref * O.float()).view(B, L, H, 64).sum(-1).permute(0, 2, 1)
    assert float((delta - want).abs().max()) < 2e-2 * float(want.abs().max())


@pytest.mark.parametrize("M", [1576, 392])
@pytest.mark.parametrize("act", [1, 2])
def test_forward_stores_activation_derivative_and_backward_multiplies(M, act):
    """c_fc forward with C2 := act'(pre-activation) (QuickGELU: from the same sigmoid as the activation) and the c_proj dgrad
    that multiplies with it (mul_aux_act = ACT_DERIV) -- against autograd; 2-CTA (M >= 512) and 1-CTA kernels."""
    from segclip_b200 import ops
    torch.manual_seed(M + act)
    N, K = 3072, 768
    x = torch.randn(M, K, device="cuda").bfloat16()
    W1 = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b1 = torch.randn(N, device="cuda") * 0.1
    hact = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    hder = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, W1, hact, bias=b1, act=act, C2=hder, c2_is_act_grad=True)
    pre = (x.float() @ W1.float().t() + b1).requires_grad_(True)
    y = pre * torch.sigmoid(1.702 * pre) if act == 1 else torch.nn.functional.gelu(pre)
    y.backward(torch.ones_like(y))
    assert float((hact.float() - y.detach()).abs().max() / y.abs().max()) < 1e-2
    assert float((hder.float() - pre.grad).abs().max()) < 1.2e-2          # derivative values are O(1)
    # backward: d_a = (dY W2) * stored derivative (+ fused column sums = c_fc bias gradient)
    dy = torch.randn(M, K, device="cuda").bfloat16()
    W2 = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    d_a = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(N, device="cuda")
    ops.gemm(dy, W2, d_a, trans_b=True, mul_aux=hder, mul_aux_act=ops.ACT_DERIV, colsum_out=cs)
    torch.cuda.synchronize()
    want = (dy.float() @ W2.float()) * hder.float()
    assert float((d_a.float() - want).abs().max() / want.abs().max()) < 1e-2
    ref = d_a.float().sum(0)
    assert float((cs - ref).abs().max()) < 2e-3 * (1 + float(ref.abs().max()))


@pytest.mark.parametrize("M", [19712, 19712 + 13])
def test_tc2_tail_column_slices(M):
    """Incomplete last wave of the 2-CTA kernel run as 64-/128-column slices (text tower: 77 row tiles x N / 256 column tiles
    on 74 CTA pairs): every hot epilogue kind, K-major and MN-major B, ragged last row tile."""
    from segclip_b200 import _lib
    if torch.cuda.get_device_properties(0).multi_processor_count != 148:
        pytest.skip("shapes chosen for 74 CTA pairs")
    k0 = _lib.kernel_launches()
    bf = torch.bfloat16
    _run(M, 512, 512, bf, bias=True, residual=True)                                        # out_proj: fp32 residual tile via TMA
    _run(M, 512, 2048, bf, bias=True, residual=True, c_dtype=bf, res_dtype=bf)             # bf16 residual stream
    _run(M, 512, 2048, bf, tb=True, c_dtype=bf)                                            # dgrad, MN-major B
    _run(M, 512, 1536, bf, tb=True, c_dtype=bf, seed=3)
    _run(M, 2048, 512, bf, bias=True, act=1, c_dtype=bf, c2=bf)                            # c_fc: two outputs, 24-tile tail
    _run(M, 1536, 512, bf, bias=True, c_dtype=bf)                                          # qkv: 18-tile tail
    _run(M, 512, 512, bf)                                                                  # plain fp32 output
    _run(M, 512, 512, bf, tb=True, accumulate=True)                                        # fp32 C +=
    k1 = _lib.kernel_launches()
    assert k1["gemm_tc2_tail"] - k0["gemm_tc2_tail"] == 8, (k0, k1)


def test_tc2_tail_slices_activation_gradient_with_colsum():
    """Text c_proj dgrad: activation-derivative operand tile + fused bias-gradient column sums, with a sliced tail."""
    from segclip_b200 import _lib, ops
    if torch.cuda.get_device_properties(0).multi_processor_count != 148:
        pytest.skip("shapes chosen for 74 CTA pairs")
    torch.manual_seed(0)
    M, N, K = 19712, 2048, 512
    dy = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    deriv = torch.rand(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(N, device="cuda")
    k0 = _lib.kernel_launches()
    ops.gemm(dy, W, out, trans_b=True, mul_aux=deriv, mul_aux_act=ops.ACT_DERIV, colsum_out=cs)
    torch.cuda.synchronize()
    k1 = _lib.kernel_launches()
    assert k1["gemm_tc2_tail"] - k0["gemm_tc2_tail"] == 1
    want = (dy.float() @ W.float()) * deriv.float()
    assert float((out.float() - want).abs().max() / want.abs().max()) < 1e-2
    ref = out.float().sum(0)
    assert float((cs - ref).abs().max()) < 2e-3 * (1 + float(ref.abs().max()))
