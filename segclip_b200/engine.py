"""Static execution plan for the SegCLIP training hot path on one B200.

For a given per-GPU batch size the engine allocates every activation / gradient buffer once and
records two flat lists of prepared native calls (``ops.Op``): the forward tape and the backward tape.
A training step replays them on the current CUDA stream -- no allocation, no host synchronisation,
no PyTorch arithmetic.  The structure follows the reference call stack

  SegCLIP.forward                    modules/modeling.py:174-256
   encode_text                       modules/module_clip.py:105-143
   encode_image / VisualTransformer  modules/module_clip.py:81-103, module_clip_vtransformer.py:55-80
   SegViT / SemanticLearnerModule    modules/module_seg_vit.py:277-314,403-452
   MAEDecoder.forward_vis            modules/module_mae.py:304-330

Canonical layout: every token tensor is row-major [batch*tokens, features] (sample-major), the
residual stream and its gradient are fp32, GEMM operands are in the compute dtype (fp32 "parity"
mode or bf16).  Gradients that are shared by several consumers live in zero-initialised fp32 buffers
and are accumulated.
"""
import os

import torch

from . import _lib as L
from . import ops

G = 8            # centres (module_seg_vit.py:349)
TAU = 0.9        # module_seg_vit.py:305
DEC_HEADS = 8    # modeling.py:147
DEC_DEPTH = 3    # modeling.py:151

CLIP_BLOCK = dict(ln1="ln_1", wqkv="attn.in_proj_weight", bqkv="attn.in_proj_bias", wo="attn.out_proj.weight",
                  bo="attn.out_proj.bias", ln2="ln_2", w1="mlp.c_fc.weight", b1="mlp.c_fc.bias",
                  w2="mlp.c_proj.weight", b2="mlp.c_proj.bias")
MAE_BLOCK = dict(ln1="norm1", wqkv="attn.qkv.weight", bqkv="attn.qkv.bias", wo="attn.proj.weight",
                 bo="attn.proj.bias", ln2="norm2", w1="mlp.fc1.weight", b1="mlp.fc1.bias",
                 w2="mlp.fc2.weight", b2="mlp.fc2.bias")

# parameters the reference recipe freezes (main_task_align.py:389-441): no gradient is
# produced for them by this engine.
FROZEN_STEM = ("clip.visual.class_embedding", "clip.visual.positional_embedding", "clip.visual.conv1.weight",
               "clip.visual.ln_pre.weight", "clip.visual.ln_pre.bias", "clip.positional_embedding",
               "clip.token_embedding.weight", "vis_mae_decoder.decoder_pos_embed")
# ... of which the reference can train everything but the fixed sin-cos decoder table (requires_grad=False in the reference
# itself, module_mae.py): an engine built with train_stem=True also produces the gradients of these (the reference computes
# them whenever the recipe does not freeze the parameters)
TRAINABLE_STEM = FROZEN_STEM[:-1]


def _is_gemm_weight(name, t):
    return t.dim() >= 2 and not any(s in name for s in (
        "positional_embedding", "token_embedding", "semantic_center", "k_conv", "v_conv", "a_fc", "mask_token",
        "decoder_pos_embed"))


class Plan:
    """Buffers + forward/backward tapes for one batch size."""

    def __init__(self):
        self.bufs = {}
        self.zero = []          # tensors cleared at the start of every step
        self.fwd = []
        self.eval_tail = {"text": [], "vision": []}     # inference-only ops (hidden states of all tokens), see infer()
        self.bwd_groups = []
        self.bwd = []
        self.slot = None        # exchange buffers of this batch size (world > 1)

    def f(self, op):
        self.fwd.append(op)

    def b(self, group):
        group = list(group)
        self.bwd_groups.append(group)
        return group

    def finish(self):
        self.bwd = [op for grp in reversed(self.bwd_groups) for op in grp]


class Engine:
    def __init__(self, cfg, named_params, precision="bf16", kv_layout="torch18_flat", rank=0, world=1, train_stem=False):
        assert precision in ("fp32", "bf16")
        self.train_stem = bool(train_stem)
        assert kv_layout in ("torch18_flat", "per_sample")
        self.cfg = dict(cfg)
        self.precision = precision
        self.T = torch.float32 if precision == "fp32" else torch.bfloat16
        self.kv_layout = kv_layout
        self.rank, self.world = rank, world
        self.params = dict(named_params)          # name -> nn.Parameter (fp32 masters)
        self.dev = next(iter(self.params.values())).device
        c = self.cfg
        self.vw, self.tw, self.E = c["vision_width"], c["text_width"], c["embed_dim"]
        self.patch, self.grid = c["patch"], c["grid"]
        self.grid_hw = (self.grid, self.grid)        # token grid of the input (inference clones may differ, see for_grid)
        self.pos_table = None                        # resized positional table of an inference clone
        self.Lp = self.grid ** 2
        self.Tctx = c["context"]
        self.Hv, self.Ht = self.vw // 64, self.tw // 64
        self.fsl = c["first_stage_layer"]
        self.n2 = 12 - self.fsl
        self.dd = self.vw // 2
        # kept tokens of the masked pass, CLS included: int(L * (1 - mask_ratio)) like module_clip_util.py:98
        self.keep = int((self.Lp + 1) * (1 - c.get("mae_vis_mask_ratio", 0.75)))
        self.Lm = self.keep - 1
        self.use_mae, self.use_kl = bool(c["use_mae"]), bool(c["use_kl"])
        self.text_layers = c["text_layers"]
        self.plans = {}
        self.eval_plans = {}
        self._ptr_sig = None
        self._setup_params()
        self.gather = None        # multi-GPU embedding exchange (segclip_b200.p2p), set by the module
        self.sync_group = None    # native gradient all-reduce (enable_grad_sync)
        self.sync_world = 1
        self.nvls = None          # NVSwitch-multicast transport of the gradient buckets (segclip_b200.allreduce)

    # ------------------------------------------------------------------ parameters
    def _setup_params(self):
        dev = self.dev
        self._layout_grads([n for n in self.params if n not in FROZEN_STEM or (self.train_stem and n in TRAINABLE_STEM)])
        self._sig()
        # compute-dtype shadows of the GEMM weights (bf16 mode); fp32 mode reads the masters directly
        self.shadow = {}
        self.cast_op = None
        if self.T != torch.float32:
            items, blocks = [], 0
            wnames = [n for n, p in self.params.items() if _is_gemm_weight(n, p)]
            tot = sum((self.params[n].numel() + 7) // 8 * 8 for n in wnames)
            self.wflat = torch.empty(tot, device=dev, dtype=self.T)
            o = 0
            host = []
            for n in wnames:
                p = self.params[n]
                self.shadow[n] = self.wflat[o:o + p.numel()].view(p.shape)
                nb = (p.numel() + 1023) // 1024
                host.append((p.data_ptr(), self.shadow[n].data_ptr(), p.numel(), blocks))
                blocks += nb
                o += (p.numel() + 7) // 8 * 8
            arr = (L.CastItem * len(host))()
            for i, (s, d, n_, fb) in enumerate(host):
                arr[i].src, arr[i].dst, arr[i].n, arr[i].first_block = s, d, n_, fb
            import ctypes
            raw = bytes(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr)))
            self.cast_items = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            self.cast_op = ops.cast_multi_op(self.cast_items, len(host), blocks, L.dt(self.wflat))
        # dense block-diagonal forms of the grouped 1x1 convs (module_seg_vit.py:266-269)
        s = "clip.visual.transformer.semantic_layer2."
        self.kdense = torch.empty(self.vw, self.vw, device=dev, dtype=self.T)
        self.vdense = torch.empty(self.vw, self.vw, device=dev, dtype=self.T)
        self.dkdense = torch.zeros(self.vw, self.vw, device=dev, dtype=torch.float32)
        self.dvdense = torch.zeros(self.vw, self.vw, device=dev, dtype=torch.float32)
        self.prep_ops = [ops.blockdiag_expand_op(self.params[s + "k_conv.weight"].data, self.kdense, self.Hv),
                         ops.blockdiag_expand_op(self.params[s + "v_conv.weight"].data, self.vdense, self.Hv)]

    def _layout_grads(self, names):
        """One flat fp32 gradient buffer; `names` fixes the order (backward-completion order when the native
        gradient all-reduce is enabled, so finished prefixes can be reduced while backward continues)."""
        self.grad_names = list(names)
        sizes = [self.params[n].numel() for n in names]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += (s + 3) // 4 * 4          # keep every view 16-byte aligned
        if getattr(self, "gflat", None) is None or self.gflat.numel() != tot:
            self.gflat = None
            if getattr(self, "nvls", None) is not None:          # symmetric memory bound to an NVSwitch multicast object
                self.gflat = self.nvls.alloc(tot)
                if self.gflat is None:
                    self.nvls = None
            if self.gflat is None:
                self.gflat = torch.zeros(tot, device=self.dev, dtype=torch.float32)
        self.goffs = dict(zip(names, offs))
        self.gsizes = dict(zip(names, sizes))
        self.grads = {n: self.gflat[o:o + s].view(self.params[n].shape) for n, o, s in zip(names, offs, sizes)}

    # ------------------------------------------------------------------ inference at another input size
    def for_grid(self, gh, gw):
        """Inference-only view of this engine for inputs of gh x gw patches (module_clip_vtransformer.py:35-53): shares every
        parameter / shadow / gradient buffer, owns its plans and a bicubically resized positional table.  The reference's
        SegViT takes its semantic branch only for n or 4 n tokens (module_seg_vit.py:423); the same rule is enforced."""
        key = (gh, gw)
        clones = self.__dict__.setdefault("_grid_clones", {})
        if key not in clones:
            n = self.grid * self.grid
            if gh * gw not in (n, 4 * n):
                raise L.SegclipB200Error("inference input of %d x %d patches: the reference's SegViT supports %d or %d patch tokens only "
                                         "(modules/module_seg_vit.py:423)" % (gh, gw, n, 4 * n))
            import copy
            c = copy.copy(self)
            c.grid_hw, c.Lp = key, gh * gw
            c.plans, c.eval_plans = {}, {}
            c.use_mae = c.use_kl = False
            c.gather, c.sync_group, c.nvls = None, None, None
            c.world, c.rank = 1, 0
            c.pos_table = torch.empty(gh * gw, self.vw, device=self.dev)
            c.pos_op = ops.bicubic_resize_op(self.P("clip.visual.positional_embedding")[1:], c.pos_table, (self.grid, self.grid), key)
            c.prep_ops = list(self.prep_ops) + [c.pos_op]          # recomputed per call: follows parameter updates
            clones[key] = c
        return clones[key]

    # ------------------------------------------------------------------ native gradient all-reduce
    def enable_grad_sync(self, group, bucket_mb=None):
        """Bucketed NCCL all-reduce of the flat gradient buffer, launched from inside the backward tape as soon as a
        prefix of the buffer is final (replaces DDP's hooks, main_task_align.py:251-252: the native backward is a single
        autograd node, so DDP could only reduce after it).  The mean (1/world) is folded into the final hand-over."""
        assert not self.plans, "enable_grad_sync must be called before the first forward"
        import torch.distributed as dist
        self.sync_group = group
        self.sync_world = dist.get_world_size(group)
        # dry run on a tiny batch: find, for every parameter, the last backward op that writes its gradient
        probe = self._build(2, probe=True)
        last = {n: -1 for n in self.grad_names}
        last.update(self._last_writers(probe))
        del probe
        order = sorted(self.grad_names, key=lambda n: (last[n], self.goffs[n]))
        self._grad_last_op = last
        # transport of the bucket reductions: the library's NVLS multimem kernel when the GPUs share an NVSwitch multicast
        # domain, else (or with SEGCLIP_GRAD_SYNC=nccl: the comparator) NCCL all-reduce
        from .allreduce import NvlsGradSync
        self.nvls = NvlsGradSync.create(group, self.dev) if self.sync_world > 1 else None
        if self.nvls is not None:
            self.gflat = None          # re-allocate as symmetric memory
        self._layout_grads(order)
        # buckets over the new order
        bucket_mb = bucket_mb or float(os.environ.get("SEGCLIP_BUCKET_MB", "48"))
        limit = int(bucket_mb * (1 << 20)) // 4
        self.buckets = []          # (start, end, name of last param)
        start = 0
        for n in order:
            end = self.goffs[n] + (self.gsizes[n] + 3) // 4 * 4
            if end - start >= limit:
                self.buckets.append((start, end, n))
                start = end
        if start < self.gflat.numel():
            self.buckets.append((start, self.gflat.numel(), order[-1]))
        self.plans = {}
        self.eval_plans = {}

    def _sig(self):
        self._ptr_sig = tuple(p.data_ptr() for p in self.params.values())

    def params_moved(self):
        return self._ptr_sig != tuple(p.data_ptr() for p in self.params.values())

    def P(self, name):
        return self.params[name].data

    def W(self, name):
        """GEMM view [out, in] of a weight in the compute dtype."""
        t = self.shadow[name] if name in self.shadow else self.params[name].data
        return t.reshape(t.shape[0], -1)

    def conv_weight(self):
        """conv1 weight as a [width, 3*p*p (padded to a multiple of 8)] GEMM operand in the compute dtype."""
        w = self.W("clip.visual.conv1.weight")
        K = w.shape[1]
        if K % 8 == 0:
            return w
        if getattr(self, "_conv_pad", None) is None:
            self._conv_pad = torch.zeros(w.shape[0], (K + 7) // 8 * 8, device=self.dev, dtype=self.T)
        return self._conv_pad

    def _refresh_conv_pad(self):
        if getattr(self, "_conv_pad", None) is not None:
            w = self.W("clip.visual.conv1.weight")
            self._conv_pad[:, :w.shape[1]].copy_(w)      # frozen stem weight: a strided copy, no arithmetic

    def Gr(self, name):
        g = self.grads[name]
        return g.reshape(g.shape[0], -1) if g.dim() >= 2 else g

    # ------------------------------------------------------------------ plan construction
    def plan(self, B):
        if B not in self.plans:
            self.plans[B] = self._build(B)
        return self.plans[B]

    def eval_plan(self, B):
        """Plan for inference at batch B.  A training plan of the same batch is reused when it exists; otherwise an
        exchange-free plan is built (local t_all / v_all, no gradient buckets): evaluation may run on one rank only
        (main_task_align.py:484-490), so it must neither enter a collective nor re-point the exchange's buffers."""
        if B in self.plans:
            return self.plans[B]
        if B not in self.eval_plans:
            self.eval_plans[B] = self._build(B, probe=True)
        return self.eval_plans[B]

    def _build(self, B, probe=False):
        pl = Plan()
        dev, T = self.dev, self.T
        f32, i32 = torch.float32, torch.int32
        is_bf16 = T != f32
        res_h, res_w = self.patch * self.grid_hw[0], self.patch * self.grid_hw[1]

        def buf(name, shape, dtype=f32, zero=False):
            assert name not in pl.bufs, name
            t = torch.zeros(shape, device=dev, dtype=dtype) if zero else torch.empty(shape, device=dev, dtype=dtype)
            pl.bufs[name] = t
            if zero:
                pl.zero.append(t)
            return t

        scratch = {}

        def sbuf(name, shape, dtype):
            key = (name, tuple(shape), dtype)
            if key not in scratch:
                scratch[key] = torch.empty(shape, device=dev, dtype=dtype)
            return scratch[key]

        def tcopy(name, x32):
            """T-typed twin of an fp32 gradient buffer (aliases it in fp32 mode)."""
            return buf(name, x32.shape, T) if is_bf16 else x32

        def sync_T(x32, xT):
            return [ops.convert_op(x32, xT)] if is_bf16 else []

        # ---------------- inputs
        ids = buf("in.ids", (B, self.Tctx), torch.int64)
        image = buf("in.image", (B, 3, res_h, res_w))
        seg = buf("in.seg", (B, self.Lp), torch.int64)
        u1 = buf("in.u1", (B, G, self.Lp))
        u2 = buf("in.u2", (B, self.Lp + 1))
        u3 = buf("in.u3", (B, G, max(self.Lm, 1)))
        forced_main = buf("in.forced_main", (B, self.Lp), i32)
        forced_mae = buf("in.forced_mae", (B, max(self.Lm, 1)), i32)
        pl.use_forced = False
        loss = buf("loss", (1,), zero=True)
        pl.zero += [self.dkdense, self.dvdense]
        center_idx = torch.arange(G, device=dev, dtype=i32).repeat(B)

        # ---------------- generic layers
        def linear_bwd(dy, x, wname, bname, dx=None, dx_acc=False, wslice=None, bslice=None, mul_aux=None, mul_act=0,
                       dx_colsum=None, simt=False, dot=None):
            w = self.W(wname)
            gw = self.Gr(wname)
            gb = self.grads[bname] if bname else None
            if wslice is not None:
                w, gw = w[wslice], gw[wslice]
            if bslice is not None:
                gb = gb[bslice]
            grp = []
            if dx is not None:
                dkw = dict(dot_aux=dot[0], dot_out=dot[1], dot_L=dot[2]) if dot is not None else {}
                grp.append(ops.gemm_op(dy, w, dx, trans_b=True, accumulate=dx_acc, mul_aux=mul_aux, mul_aux_act=mul_act,
                                       colsum_out=dx_colsum, force_simt=simt, **dkw))
            grp.append(ops.gemm_op(dy, x, gw, trans_a=True, trans_b=True, accumulate=True, split_k=0 if simt else -1,
                                   force_simt=simt))
            if gb is not None:
                grp.append(ops.colsum_op(dy, gb))
            return grp

        def ln_fwd(x, pre, y, tag, eps=1e-5, remap=None, rows=None):
            n = x.shape[0]
            mean, rstd = buf(tag + ".mean", (n,)), buf(tag + ".rstd", (n,))
            pl.f(ops.layernorm_op(x, self.P(pre + ".weight"), self.P(pre + ".bias"), y, eps, mean, rstd, remap))
            return mean, rstd

        def ln_bwd(dy, x, stats, pre, dx=None, acc=False, dx_copy=None, remap=None, colsum=None):
            return ops.layernorm_bwd_op(dy, x, stats[0], stats[1], self.P(pre + ".weight"), dx, acc,
                                        dx_copy if (is_bf16 and dx_copy is not dx) else None,
                                        self.grads[pre + ".weight"], self.grads[pre + ".bias"],
                                        remap, colsum)

        def mlp(x_mid, x_out, dx, dxT, pre, nm, tag, eps, act, bo_grad=None):
            """x_out = x_mid + W2 act(W1 LN2(x_mid) + b1) + b2; backward leaves d x_mid in dx/dxT."""
            M, D = x_mid.shape
            Dh = self.params[pre + nm["w1"]].shape[0]
            h2 = buf(tag + ".ln2", (M, D), T)
            st2 = ln_fwd(x_mid, pre + nm["ln2"], h2, tag + ".ln2", eps)
            hpre, hact = buf(tag + ".fc_pre", (M, Dh), T), buf(tag + ".fc_act", (M, Dh), T)
            # bf16 mode: the forward epilogue stores act'(pre-activation) instead of the pre-activation (it has the sigmoid /
            # erf at hand); the fused activation backward in the c_proj dgrad epilogue is then a plain multiply
            deriv = is_bf16 and not os.environ.get("SC_NO_FUSE_ACT") and not os.environ.get("SC_NO_ACT_DERIV")
            pl.f(ops.gemm_op(h2, self.W(pre + nm["w1"]), hact, bias=self.P(pre + nm["b1"]), act=act, C2=hpre, c2_is_act_grad=deriv))
            pl.f(ops.gemm_op(hact, self.W(pre + nm["w2"]), x_out, bias=self.P(pre + nm["b2"]), residual=x_mid))
            d_a, d_ln = sbuf("d_a", (M, Dh), T), sbuf("d_ln", (M, D), T)
            if os.environ.get("SC_NO_FUSE_ACT"):
                grp = linear_bwd(dxT, hact, pre + nm["w2"], pre + nm["b2"], dx=d_a)
                grp.append(ops.act_bwd_op(d_a, hpre, d_a, act))
            else:
                grp = linear_bwd(dxT, hact, pre + nm["w2"], pre + nm["b2"], dx=d_a, mul_aux=hpre, mul_act=ops.ACT_DERIV if deriv else act,  # fused act'()
                                 dx_colsum=self.grads[pre + nm["b1"]] if is_bf16 else None)                    # + c_fc bias grad
            fused_b1 = is_bf16 and not os.environ.get("SC_NO_FUSE_ACT")
            grp += linear_bwd(d_a, h2, pre + nm["w1"], None if fused_b1 else pre + nm["b1"], dx=d_ln)
            grp.append(ln_bwd(d_ln, x_mid, st2, pre + nm["ln2"], dx, True, dxT, colsum=bo_grad))
            return grp

        def block(x_in, dx, dxT, pre, nm, tag, Bn, Lseq, H, causal=False, eps=1e-5, act=ops.ACT_QUICKGELU, prev=None):
            """Pre-LN residual block (module_seg_vit.py:191-196, module_clip_ttransformer.py:48-52,
            module_mae.py:199-201).  The forward residual stream is fp32 (like the reference and like autocast: rounding it
            to bf16 after every add was emulated on the oracle -- bias / LayerNorm gradients move by 0.25-0.35 rel-L2, three
            times the operand-rounding floor).  dx/dxT: gradient of the stream, two layouts:
              * dx is dxT -- ONE buffer in the compute dtype (the big streams: text, layers0 main / MAE pass): LayerNorm
                backward accumulates into it in place; the same emulation shows no measurable cost (max rel-L2 unchanged);
              * dx fp32 + dxT compute-dtype twin (small streams fed by fp32-only glue kernels)."""
            M, D = x_in.shape
            sdt = f32
            hd = D // H
            h1 = buf(tag + ".ln1", (M, D), T)
            st1 = ln_fwd(x_in, pre + nm["ln1"], h1, tag + ".ln1", eps)
            qkv = buf(tag + ".qkv", (M, 3 * D), T)
            pl.f(ops.gemm_op(h1, self.W(pre + nm["wqkv"]), qkv, bias=self.P(pre + nm["bqkv"])))
            att, lse = buf(tag + ".att", (M, D), T), buf(tag + ".lse", (Bn, H, Lseq))
            strides = (Lseq * 3 * D, 3 * D)
            ad = ops.attn_desc(qkv, qkv[:, D:], qkv[:, 2 * D:], att, lse, Bn, H, Lseq, Lseq, hd, strides, strides, strides,
                               (Lseq * D, D), causal)
            pl.f(ops.attention_op(ad, (qkv, att, lse)))
            x_mid = buf(tag + ".x_mid", (M, D), sdt)
            pl.f(ops.gemm_op(att, self.W(pre + nm["wo"]), x_mid, bias=self.P(pre + nm["bo"]), residual=x_in))
            x_out = buf(tag + ".x_out", (M, D), sdt)
            # bias gradients that are column sums of a residual-stream gradient are produced by the LayerNorm backward
            # that writes that gradient: LN2-bwd -> out_proj bias of this block, LN1-bwd -> c_proj bias of the previous block
            # (the column sums ride in registers of the LayerNorm backward, sc_layernorm_bwd `dx_colsum`; the first version with
            # shared-memory accumulators cost more than the stand-alone colsum kernels it removed)
            fuse_ln = not os.environ.get("SC_NO_FUSE_LN_COLSUM")
            grp_mlp = mlp(x_mid, x_out, dx, dxT, pre, nm, tag, eps, act, bo_grad=self.grads[pre + nm["bo"]] if fuse_ln else None)
            d_att, dqkv, d_ln = sbuf("d_att", (M, D), T), sbuf("dqkv", (M, 3 * D), T), sbuf("d_ln", (M, D), T)
            # delta = rowsum(dO o O) of the attention backward comes out of the out_proj dgrad that computes dO (head dim 64)
            att_delta = sbuf("att_delta", (Bn, H, Lseq), f32)
            fuse_delta = is_bf16 and hd == 64 and not os.environ.get("SC_NO_FUSE_DELTA")
            grp = linear_bwd(dxT, att, pre + nm["wo"], None if fuse_ln else pre + nm["bo"], dx=d_att,
                             dot=(att, att_delta, Lseq) if fuse_delta else None)
            # in_proj bias gradient = column sums of dQ | dK | dV: produced by the attention backward that writes them
            fuse_qkv_bias = is_bf16 and not os.environ.get("SC_NO_FUSE_QKV_BIAS")
            grp.append(ops.attention_bwd_op(ad, d_att, dqkv, dqkv[:, D:], dqkv[:, 2 * D:], att_delta,
                                            bias_grad=self.grads[pre + nm["bqkv"]] if fuse_qkv_bias else None, delta_ready=fuse_delta))
            grp += linear_bwd(dqkv, h1, pre + nm["wqkv"], None if fuse_qkv_bias else pre + nm["bqkv"], dx=d_ln)
            prev_b2 = None
            if prev is not None and fuse_ln:          # take over the previous block's c_proj bias gradient
                prev["group"].remove(prev["colsum"])
                prev_b2 = prev["b2"]
            grp.append(ln_bwd(d_ln, x_in, st1, pre + nm["ln1"], dx, True, dxT, colsum=prev_b2))
            pl.b(grp)          # executed after the MLP group (groups run in reverse order)
            stored = pl.b(grp_mlp)
            b2_colsum = [op for op in stored if op.name == "sc_colsum"][0]      # colsum(dxT -> c_proj bias)
            handle = dict(group=stored, colsum=b2_colsum, b2=self.grads[pre + nm["b2"]])
            return x_out, handle

        def cross_block(q_in, xp, dq, dqT, d_xp, pre, tag, Bn, Lx):
            """CrossAttentionBlock (module_seg_vit.py:213-218) incl. both K/V layouts."""
            D, H = self.vw, self.Hv
            S = G + Lx
            Mq = Bn * G
            qn = buf(tag + ".qn", (Mq, D), T)
            st_x = ln_fwd(q_in, pre + "ln_x", qn, tag + ".ln_x")
            kvn = buf(tag + ".kvn", (Bn * S, D), T)
            st_kq = ln_fwd(q_in, pre + "ln_k", kvn, tag + ".ln_kq", remap=(G, S, 0))
            st_kx = ln_fwd(xp, pre + "ln_k", kvn, tag + ".ln_kx", remap=(Lx, S, G))
            w, bias = self.W(pre + "attn.in_proj_weight"), self.P(pre + "attn.in_proj_bias")
            qp, kvp = buf(tag + ".qp", (Mq, D), T), buf(tag + ".kvp", (Bn * S, 2 * D), T)
            pl.f(ops.gemm_op(qn, w[:D], qp, bias=bias[:D]))
            pl.f(ops.gemm_op(kvn, w[D:], kvp, bias=bias[D:]))
            o, lse = buf(tag + ".o", (Mq, D), T), buf(tag + ".lse", (Bn, H, G))
            kstr = (2 * D, Bn * 2 * D) if self.kv_layout == "torch18_flat" else (S * 2 * D, 2 * D)
            ad = ops.attn_desc(qp, kvp, kvp[:, D:], o, lse, Bn, H, G, S, D // H, (G * D, D), kstr, kstr, (G * D, D))
            pl.f(ops.attention_op(ad, (qp, kvp, o, lse)))
            q_mid = buf(tag + ".q_mid", (Mq, D))
            pl.f(ops.gemm_op(o, self.W(pre + "attn.out_proj.weight"), q_mid, bias=self.P(pre + "attn.out_proj.bias"),
                             residual=q_in))
            q_out = buf(tag + ".q_out", (Mq, D))
            grp_mlp = mlp(q_mid, q_out, dq, dqT, pre, CLIP_BLOCK, tag, 1e-5, ops.ACT_QUICKGELU)
            d_o, dqp = sbuf("d_att", (Mq, D), T), sbuf("dqp", (Mq, D), T)
            dkvp = sbuf("dkvp", (Bn * S, 2 * D), T)
            d_qn, d_kvn = sbuf("d_ln", (Mq, D), T), sbuf("d_kvn", (Bn * S, D), T)
            grp = linear_bwd(dqT, o, pre + "attn.out_proj.weight", pre + "attn.out_proj.bias", dx=d_o)
            grp.append(ops.attention_bwd_op(ad, d_o, dqp, dkvp, dkvp[:, D:], sbuf("att_delta_x", (Bn, H, G), f32)))
            grp += linear_bwd(dqp, qn, pre + "attn.in_proj_weight", pre + "attn.in_proj_bias", dx=d_qn,
                              wslice=slice(0, D), bslice=slice(0, D))
            grp += linear_bwd(dkvp, kvn, pre + "attn.in_proj_weight", pre + "attn.in_proj_bias", dx=d_kvn,
                              wslice=slice(D, 3 * D), bslice=slice(D, 3 * D))
            grp.append(ln_bwd(d_kvn, xp, st_kx, pre + "ln_k", d_xp, True, None, remap=(Lx, S, G)))
            grp.append(ln_bwd(d_qn, q_in, st_x, pre + "ln_x", dq, True, None))
            grp.append(ln_bwd(d_kvn, q_in, st_kq, pre + "ln_k", dq, True, dqT, remap=(G, S, 0)))
            pl.b(grp)
            pl.b(grp_mlp)
            return q_out

        def semantic(xp, d_xp, u, forced, d_hard_extra, tag, Bn, Lx):
            """SemanticLearnerModule.forward (module_seg_vit.py:277-314).  Returns (sx, d_sx, idx)."""
            s = "clip.visual.transformer.semantic_layer2."
            D = self.vw
            Mq, Mx = Bn * G, Bn * Lx
            xin = buf(tag + ".xin", (Mx, D), T)
            st_norm = ln_fwd(xp, s + "norm", xin, tag + ".norm")
            q0 = buf(tag + ".q0", (Mq, D))
            pl.f(ops.gather_rows_op(self.P(s + "semantic_center"), center_idx, q0))
            dq, dqT = buf(tag + ".dq", (Mq, D)), None
            dqT = tcopy(tag + ".dqT", dq)
            # backward of q0 = repeat(semantic_center): sum over the batch
            pl.b([ops.colsum_op(dq, self.grads[s + "semantic_center"].view(-1), rows=Bn, cols=G * D, ld=G * D)])
            q1 = cross_block(q0, xp, dq, dqT, d_xp, s + "cross_att.0.", tag + ".ca0", Bn, Lx)
            q2 = cross_block(q1, xp, dq, dqT, d_xp, s + "cross_att.1.", tag + ".ca1", Bn, Lx)
            qf = buf(tag + ".qf", (Mq, D))
            st_cross = ln_fwd(q2, s + "cross_ln", qf, tag + ".cross_ln")
            kfeat, vfeat = buf(tag + ".kfeat", (Mx, D)), buf(tag + ".vfeat", (Mx, D), T)
            pl.f(ops.gemm_op(xin, self.kdense, kfeat))
            pl.f(ops.gemm_op(xin, self.vdense, vfeat))
            k = buf(tag + ".k", (Mx, D))
            st_kln = ln_fwd(kfeat, s + "k_ln", k, tag + ".k_ln")
            y_soft, soft = buf(tag + ".y_soft", (Bn, G, Lx)), buf(tag + ".soft", (Bn, G, Lx))
            idx, count = buf(tag + ".idx", (Bn, Lx), i32), buf(tag + ".count", (Bn, G), zero=True)
            agg, ssum = buf(tag + ".agg", (Mq, D)), buf(tag + ".sum", (Mq, D))
            # assignment softmax + hard arg-max + per-centre weighted mean: ONE kernel, one CTA per sample (aggregate.cu)
            fop = ops.assign_aggregate_fwd_op(qf, k, u, y_soft, idx, count, vfeat, agg, ssum, Bn, Lx, D, TAU, None, soft)
            fop_forced = ops.assign_aggregate_fwd_op(qf, k, u, y_soft, idx, count, vfeat, agg, ssum, Bn, Lx, D, TAU, forced, soft)
            fop_eval = ops.assign_aggregate_fwd_op(qf, k, None, y_soft, idx, count, vfeat, agg, ssum, Bn, Lx, D, TAU, None, soft)
            pl.f((fop, fop_forced, fop_eval))          # (training, teacher-forced, inference) variants
            # proj_o = LN -> fc1 -> erf-GELU -> fc2 -> QuickGELU (module_seg_vit.py:271-275)
            hp = buf(tag + ".po_ln", (Mq, D), T)
            st_po = ln_fwd(ssum, s + "proj_o.ln", hp, tag + ".po_ln")
            pre1, a1 = buf(tag + ".po_pre1", (Mq, 4 * D), T), buf(tag + ".po_a1", (Mq, 4 * D), T)
            pl.f(ops.gemm_op(hp, self.W(s + "proj_o.mlp.fc1.weight"), a1, bias=self.P(s + "proj_o.mlp.fc1.bias"),
                             act=ops.ACT_GELU_ERF, C2=pre1))
            pre2, sx = buf(tag + ".po_pre2", (Mq, D)), buf(tag + ".sx", (Mq, D))
            pl.f(ops.gemm_op(a1, self.W(s + "proj_o.mlp.fc2.weight"), sx, bias=self.P(s + "proj_o.mlp.fc2.bias"),
                             act=ops.ACT_QUICKGELU, C2=pre2))
            # ---- backward (execution order)
            d_sx = buf(tag + ".d_sx", (Mq, D))
            d_pre2, d_a1, d_h = sbuf("d_pre2", (Mq, D), T), sbuf("d_a", (Mq, 4 * D), T), sbuf("d_ln", (Mq, D), T)
            d_sum, d_qf = buf(tag + ".d_sum", (Mq, D)), buf(tag + ".d_qf", (Mq, D))
            d_logits = buf(tag + ".d_logits", (Bn, G, Lx))
            d_v, d_k = sbuf("d_v", (Mx, D), T), sbuf("d_k", (Mx, D), f32)
            d_kfeat, d_xin = sbuf("d_kfeat", (Mx, D), T), sbuf("d_xin", (Mx, D), f32)
            grp = [ops.act_bwd_op(d_sx, pre2, d_pre2, ops.ACT_QUICKGELU)]
            grp += linear_bwd(d_pre2, a1, s + "proj_o.mlp.fc2.weight", s + "proj_o.mlp.fc2.bias", dx=d_a1, mul_aux=pre1,
                              mul_act=ops.ACT_GELU_ERF)
            grp += linear_bwd(d_a1, hp, s + "proj_o.mlp.fc1.weight", s + "proj_o.mlp.fc1.bias", dx=d_h)
            grp.append(ln_bwd(d_h, ssum, st_po, s + "proj_o.ln", d_sum))
            grp.append(ops.assign_bwd_op(d_sum, agg, vfeat, idx, count, y_soft, d_hard_extra, qf, k, d_logits, d_v, d_k,
                                         d_sum, d_qf, Bn, Lx, D, TAU))
            grp.append(ln_bwd(d_k, kfeat, st_kln, s + "k_ln", d_kfeat))
            grp.append(ops.gemm_op(d_kfeat, self.kdense, d_xin, trans_b=True))
            grp.append(ops.gemm_op(d_v, self.vdense, d_xin, trans_b=True, accumulate=True))
            grp.append(ops.gemm_op(d_kfeat, xin, self.dkdense, trans_a=True, trans_b=True, accumulate=True, split_k=-1))
            grp.append(ops.gemm_op(d_v, xin, self.dvdense, trans_a=True, trans_b=True, accumulate=True, split_k=-1))
            grp.append(ln_bwd(d_xin, xp, st_norm, s + "norm", d_xp, False))      # first writer of d_xp (this group runs first)
            grp.append(ln_bwd(d_qf, q2, st_cross, s + "cross_ln", dq, False, dqT))
            pl.b(grp)
            return sx, d_sx, idx

        def vision_stem(tag, rows_per_img, patch_idx, d_x0=None):
            """conv1 (as GEMM) + positional embedding + ln_pre on patch tokens only: the CLS token is
            discarded by SegViT (module_seg_vit.py:419), so class_embedding and positional row 0 receive a zero gradient.
            The reference recipe freezes all of this (F10); with train_stem the backward of the stem is appended:
            d_x0 = gradient of the stem output (the stream gradient after block 0's backward)."""
            v = "clip.visual."
            M = B * rows_per_img
            K = 3 * self.patch * self.patch
            Kp = (K + 7) // 8 * 8          # TMA needs 16-byte row pitches: patch 14 -> 588 -> 592 (zero padded)
            cols = buf(tag + ".cols", (M, Kp), T, zero=(Kp != K))
            if Kp != K:
                pl.zero[:] = [z for z in pl.zero if z is not cols]      # padding zeroed once; im2col never touches it
            pl.f(ops.im2col_op(image, cols, patch_idx, rows_per_img, self.grid_hw, self.patch))
            pre = buf(tag + ".pre", (M, self.vw))
            pos = self.P(v + "positional_embedding")[1:] if self.pos_table is None else self.pos_table
            pl.f(ops.gemm_op(cols, self.conv_weight(), pre, rowbias=pos, rowbias_idx=patch_idx, rowbias_mod=self.Lp))
            x0 = buf(tag + ".x0", (M, self.vw))
            if not (self.train_stem and d_x0 is not None):
                pl.f(ops.layernorm_op(pre, self.P(v + "ln_pre.weight"), self.P(v + "ln_pre.bias"), x0))
                return x0
            st_pre = ln_fwd(pre, v + "ln_pre", x0, tag + ".ln_pre")
            d_pre = sbuf("stem.d_pre", (M, self.vw), T)
            grp = [ln_bwd(d_x0, pre, st_pre, v + "ln_pre", d_pre)]
            # conv1.weight [width, 3, p, p] as [width, K]: dW += d_pre^T cols (K = 588 for patch 14: the padded im2col columns
            # 588..591 are dropped by the GEMM's N)
            gw = self.grads[v + "conv1.weight"].view(self.vw, K)
            grp.append(ops.gemm_op(d_pre, cols[:, :K] if Kp != K else cols, gw, trans_a=True, trans_b=True, accumulate=True,
                                   split_k=-1))
            # positional rows 1..: sum over the batch (all patches present) / scatter by patch index (masked pass)
            gpos = self.grads[v + "positional_embedding"]
            if patch_idx is None:
                grp.append(ops.colsum_op(d_pre, gpos[1:].reshape(-1), rows=B, cols=rows_per_img * self.vw, ld=rows_per_img * self.vw))
            else:
                grp.append(ops.scatter_add_rows_op(d_pre, patch_idx, gpos, idx_offset=1))
            pl.b(grp)
            return x0

        t_ = "clip.visual.transformer."
        D = self.vw

        # =============================================================== text tower
        W_, Mt = self.tw, B * self.Tctx
        xt = buf("t.x0", (Mt, W_))
        eot = buf("t.eot", (B,), i32)
        pl.f(ops.text_embed_op(ids, self.P("clip.token_embedding.weight"), self.P("clip.positional_embedding"), xt, eot, B,
                               self.Tctx, W_))
        dxt = buf("t.dx", (Mt, W_), zero=True)          # fp32 landing buffer of the EOT-row scatter; the blocks use dxtT only
        dxtT = tcopy("t.dxT", dxt)
        if self.train_stem:
            # x0 = token_embedding[ids] + positional_embedding (module_clip.py:107-109): after block 0's backward dxtT is d x0
            pl.b([ops.colsum_op(dxtT, self.grads["clip.positional_embedding"].view(-1), rows=B, cols=self.Tctx * W_, ld=self.Tctx * W_),
                  ops.scatter_add_rows_op(dxtT, ids.view(-1), self.grads["clip.token_embedding.weight"])])
        hd_ = None
        for i in range(self.text_layers):
            xt, hd_ = block(xt, dxtT, dxtT, f"clip.transformer.resblocks.{i}.", CLIP_BLOCK, f"t{i}", B, self.Tctx, self.Ht, True,
                            prev=hd_)
        xe = buf("t.xe", (B, W_))
        pl.f(ops.gather_rows_op(xt, eot, xe))
        he = buf("t.he", (B, W_), T)
        st_f = ln_fwd(xe, "clip.ln_final", he, "t.ln_final")
        t_raw = buf("t.raw", (B, self.E))
        pl.f(ops.gemm_op(he, self.W("clip.text_projection"), t_raw, trans_b=True))
        hT, hidT = buf("t.h_all", (Mt, W_), T), buf("t.hidden_all", (Mt, self.E))
        pl.eval_tail["text"] += [ops.layernorm_op(xt, self.P("clip.ln_final.weight"), self.P("clip.ln_final.bias"), hT),
                         ops.gemm_op(hT, self.W("clip.text_projection"), hidT, trans_b=True)]
        d_traw = buf("t.d_raw", (B, self.E))
        d_trawT, d_he, d_xe = tcopy("t.d_rawT", d_traw), sbuf("t.d_he", (B, W_), T), buf("t.d_xe", (B, W_))
        grp = sync_T(d_traw, d_trawT)
        grp.append(ops.gemm_op(d_trawT, self.W("clip.text_projection"), d_he))
        grp.append(ops.gemm_op(he, d_trawT, self.Gr("clip.text_projection"), trans_a=True, trans_b=True, accumulate=True))
        grp.append(ln_bwd(d_he, xe, st_f, "clip.ln_final", d_xe))
        grp.append(ops.scatter_rows_op(d_xe, eot, dxt))
        grp += sync_T(dxt, dxtT)
        pl.b(grp)

        # =============================================================== vision tower, main pass
        pl.f("wait_image")        # the image H2D copy runs on a side stream underneath the text tower
        Mv = B * self.Lp
        dxv = buf("v.dx", (Mv, D), T)          # gradient of the layers0 stream; first written (not accumulated) by semantic()
        xv = vision_stem("v.stem", self.Lp, None, dxv)
        hd_ = None
        for i in range(self.fsl):
            xv, hd_ = block(xv, dxv, dxv, f"{t_}layers0.{i}.", CLIP_BLOCK, f"v{i}", B, self.Lp, self.Hv, prev=hd_)
        d_hard_kl = buf("v.d_hard_kl", (B, G, self.Lp)) if self.use_kl else None
        sx, d_sx, idx_main = semantic(xv, dxv, u1, forced_main, d_hard_kl, "v.sem", B, self.Lp)
        if self.use_kl:
            pl.f(ops.superpixel_kl_op(idx_main, seg, loss, d_hard_kl, B, self.Lp))
        dc, c = d_sx, sx
        dcT = tcopy("v.dcT", dc)
        pl.b(sync_T(dc, dcT) if self.n2 == 0 else [])
        hd_ = None
        for i in range(self.n2):
            c, hd_ = block(c, dc, dcT, f"{t_}layers2.{i}.", CLIP_BLOCK, f"v2_{i}", B, G, self.Hv, prev=hd_)
        pooled, parg = buf("v.pooled", (B, D)), buf("v.parg", (B, D), i32)
        pl.f(ops.pool_max_op(c, pooled, parg, B, G, D))
        pl.f("force_pool")      # test hook: teacher-forced arg-max routing of the max pooling
        hpv = buf("v.hp", (B, D), T)
        st_post = ln_fwd(pooled, "clip.visual.ln_post", hpv, "v.ln_post")
        v_raw = buf("v.raw", (B, self.E))
        pl.f(ops.gemm_op(hpv, self.W("clip.visual.proj"), v_raw, trans_b=True))
        cat9 = buf("v.cat9", (B * (G + 1), D))
        idx_cls = (torch.arange(B, device=dev, dtype=i32) * (G + 1)).contiguous()
        idx_ctr = (torch.arange(B * G, device=dev, dtype=i32) + torch.arange(B, device=dev, dtype=i32).repeat_interleave(G) + 1).contiguous()
        h9, hid9 = buf("v.h9", (B * (G + 1), D), T), buf("v.hidden9", (B * (G + 1), self.E))
        pl.eval_tail["vision"] += [ops.scatter_rows_op(pooled, idx_cls, cat9), ops.scatter_rows_op(c, idx_ctr, cat9),
                         ops.layernorm_op(cat9, self.P("clip.visual.ln_post.weight"), self.P("clip.visual.ln_post.bias"), h9),
                         ops.gemm_op(h9, self.W("clip.visual.proj"), hid9, trans_b=True)]
        d_vraw = buf("v.d_raw", (B, self.E))
        d_vrawT, d_hpv, d_pooled = tcopy("v.d_rawT", d_vraw), sbuf("v.d_hp", (B, D), T), buf("v.d_pooled", (B, D))
        grp = sync_T(d_vraw, d_vrawT)
        grp.append(ops.gemm_op(d_vrawT, self.W("clip.visual.proj"), d_hpv))
        grp.append(ops.gemm_op(hpv, d_vrawT, self.Gr("clip.visual.proj"), trans_a=True, trans_b=True, accumulate=True))
        grp.append(ln_bwd(d_hpv, pooled, st_post, "clip.visual.ln_post", d_pooled))
        grp.append(ops.pool_max_bwd_op(d_pooled, parg, dc, B, G, D))
        grp += sync_T(dc, dcT)
        pl.b(grp)

        # =============================================================== contrastive head
        # (modules/modeling.py:204-209,338-357)
        N = B * self.world
        if self.world > 1 and not probe:
            if self.gather is None:
                raise L.SegclipB200Error("world_size > 1: attach an exchange (segclip_b200.p2p.EmbeddingExchange) first")
            pl.slot = self.gather.slot(B, self.E)          # per-batch-size buffers owned by the exchange
            t_all, v_all, lse_ext = pl.slot.t_all, pl.slot.v_all, pl.slot.lse_all
            pl.bufs["c.t_all"], pl.bufs["c.v_all"] = t_all, v_all
        else:
            t_all, v_all = buf("c.t_all", (N, self.E)), buf("c.v_all", (N, self.E))
        lo = self.rank * B
        t_n, v_n = t_all[lo:lo + B], v_all[lo:lo + B]
        t_inv, v_inv = buf("c.t_inv", (B,)), buf("c.v_inv", (B,))
        pl.f(ops.l2norm_fwd_op(t_raw, t_n, t_inv))
        pl.f(ops.l2norm_fwd_op(v_raw, v_n, v_inv))
        pl.f("gather_embeddings")
        raw_t2v, raw_v2t = buf("c.t2v", (B, N)), buf("c.v2t", (B, N))
        pl.f(ops.gemm_op(t_n, v_all, raw_t2v))
        pl.f(ops.gemm_op(v_n, t_all, raw_v2t))
        if self.world > 1 and not probe:
            lse_all = lse_ext
            pl.bufs["c.lse_all"] = lse_all
        else:
            lse_all = buf("c.lse_all", (2, N))        # [0] = t2v rows, [1] = v2t rows, global order
        scale_p = self.P("clip.logit_scale")
        pl.f(ops.ce_lse_op(raw_t2v, lo, scale_p, lse_all[0, lo:lo + B], loss))
        pl.f(ops.ce_lse_op(raw_v2t, lo, scale_p, lse_all[1, lo:lo + B], loss))
        pl.f("gather_lse")
        d_tn, d_vn = buf("c.d_tn", (B, self.E)), buf("c.d_vn", (B, self.E))
        g_scale = self.grads["clip.logit_scale"].view(1)
        grp = [ops.ce_grad_op(raw_t2v, lo, scale_p, lse_all[0, lo:lo + B], lse_all[1], g_scale),
               ops.ce_grad_op(raw_v2t, lo, scale_p, lse_all[1, lo:lo + B], lse_all[0], g_scale),
               ops.gemm_op(raw_t2v, v_all, d_tn, trans_b=True),
               ops.gemm_op(raw_v2t, t_all, d_vn, trans_b=True),
               "release_exchange",          # last reader of the gathered embeddings / LSEs is done
               ops.l2norm_bwd_op(d_tn, t_n, t_inv, d_traw),
               ops.l2norm_bwd_op(d_vn, v_n, v_inv, d_vraw)]
        pl.b(grp)

        # =============================================================== vision tower, MAE pass
        if self.use_mae:
            L1, keep, Lm = self.Lp + 1, self.keep, self.Lm
            ids_restore, ids_keep = buf("m.ids_restore", (B, L1), i32), buf("m.ids_keep", (B, keep), i32)
            mask, pidx = buf("m.mask", (B, L1)), buf("m.pidx", (B * Lm,), i32)
            pl.f(ops.mae_mask_op(u2, ids_restore, ids_keep, mask, pidx, B, L1, keep))
            Mm = B * Lm
            dxm = buf("m.dx", (Mm, D), T)
            xm = vision_stem("m.stem", Lm, pidx, dxm)
            hd_ = None
            for i in range(self.fsl):
                xm, hd_ = block(xm, dxm, dxm, f"{t_}layers0.{i}.", CLIP_BLOCK, f"m{i}", B, Lm, self.Hv, prev=hd_)
            d_hard_rec = buf("m.d_hard_rec", (B, G, Lm))
            sxm, d_sxm, idx_mae = semantic(xm, dxm, u3, forced_mae, d_hard_rec, "m.sem", B, Lm)
            r = t_ + "reconstruct_layer2.rec_proj_a.a_fc."
            rpre, rout = buf("m.rec_pre", (Mm, D)), buf("m.rec_out", (Mm, D))
            pl.f(ops.reconstruct_fwd_op(sxm, idx_mae, self.P(r + "weight"), self.P(r + "bias"), rpre, rout, B, Lm, D))
            dr = buf("m.dr", (Mm, D))
            drT = tcopy("m.drT", dr)
            pl.b([ops.reconstruct_bwd_op(dr, rpre, sxm, idx_mae, self.P(r + "weight"), self.P(r + "bias"), d_sxm, d_hard_rec,
                                         self.grads[r + "weight"], self.grads[r + "bias"], B, Lm, D)])
            y = rout
            hd_ = None
            for i in range(self.n2):
                y, hd_ = block(y, dr, drT, f"{t_}layers_mae2.{i}.", CLIP_BLOCK, f"m2_{i}", B, Lm, self.Hv, prev=hd_)
            m_ = "vis_mae_decoder."
            dd = self.dd
            hc = buf("m.hc", (B * keep, D), T)
            pl.f(ops.mean_cat_op(y, hc, B, Lm, D))
            emb = buf("m.emb", (B * keep, dd), T)
            pl.f(ops.gemm_op(hc, self.W(m_ + "decoder_embed.weight"), emb, bias=self.P(m_ + "decoder_embed.bias")))
            xd = buf("m.xd", (B * L1, dd))
            pl.f(ops.mae_unshuffle_op(emb, self.P(m_ + "mask_token"), ids_restore, self.P(m_ + "decoder_pos_embed"), xd, B, L1,
                                      keep, dd))
            d_emb, d_hc = sbuf("m.d_emb", (B * keep, dd), T), sbuf("m.d_hc", (B * keep, D), T)
            dxd = buf("m.dxd", (B * L1, dd))
            dxdT = tcopy("m.dxdT", dxd)
            grp = [ops.mae_unshuffle_bwd_op(dxd, ids_restore, d_emb, self.grads[m_ + "mask_token"].view(-1), B, L1, keep, dd)]
            grp += linear_bwd(d_emb, hc, m_ + "decoder_embed.weight", m_ + "decoder_embed.bias", dx=d_hc)
            grp.append(ops.mean_cat_bwd_op(d_hc, dr, B, Lm, D))
            grp += sync_T(dr, drT)
            pl.b(grp)
            hd_ = None
            for i in range(DEC_DEPTH):
                xd, hd_ = block(xd, dxd, dxdT, f"{m_}decoder_blocks.{i}.", MAE_BLOCK, f"md{i}", B, L1, DEC_HEADS, False, 1e-6,
                                ops.ACT_GELU_ERF, prev=hd_)
            hn = buf("m.hn", (B * L1, dd), T)
            st_dn = ln_fwd(xd, m_ + "decoder_norm", hn, "m.dec_norm", 1e-6)
            Pp = 3 * self.patch * self.patch
            pred, dpred = buf("m.pred", (B * L1, Pp), T), buf("m.dpred", (B * L1, Pp), T)
            # 3*p*p output columns: not a multiple of 8 for patch 14 (588) -> this one layer runs on the exact-FMA kernel
            # (TMA needs 16-byte row pitches); known slow spot of the ViT-L/14 + MAE configuration
            pred_simt = Pp % 8 != 0
            pl.f(ops.gemm_op(hn, self.W(m_ + "decoder_pred.weight"), pred, bias=self.P(m_ + "decoder_pred.bias"),
                             force_simt=pred_simt))
            pl.f(ops.mae_loss_op(pred, image, mask, loss, dpred, B, L1, keep, self.grid, self.patch))
            d_hn = sbuf("m.d_hn", (B * L1, dd), T)
            grp = linear_bwd(dpred, hn, m_ + "decoder_pred.weight", m_ + "decoder_pred.bias", dx=d_hn, simt=pred_simt)
            grp.append(ln_bwd(d_hn, xd, st_dn, m_ + "decoder_norm", dxd, False, dxdT))
            pl.b(grp)

        # conv-weight gradients: fold the dense block-diagonal gradients back (runs last)
        s = t_ + "semantic_layer2."
        pl.bwd_groups.insert(0, [
            ops.blockdiag_reduce_op(self.dkdense, self.grads[s + "k_conv.weight"].view(self.vw, -1), self.Hv),
            ops.blockdiag_reduce_op(self.dvdense, self.grads[s + "v_conv.weight"].view(self.vw, -1), self.Hv)])
        pl.finish()
        pl.scratch = scratch
        pl.loss = loss
        pl.bucket_after = {}
        if self.sync_group is not None and not probe:
            # op index after which each bucket is final = last writer of its last-finishing parameter
            names_by_bucket = []
            for (s, e_, _) in self.buckets:
                names_by_bucket.append([n for n in self.grad_names if s <= self.goffs[n] < e_])
            lo, hi = self.gflat.data_ptr(), self.gflat.data_ptr() + self.gflat.numel() * 4
            last = self._last_writers(pl)
            for k, names in enumerate(names_by_bucket):
                idx = max([last.get(n, -1) for n in names] + [-1])
                pl.bucket_after.setdefault(idx, []).append(k)
        return pl

    def _last_writers(self, pl):
        import bisect
        lo, hi = self.gflat.data_ptr(), self.gflat.data_ptr() + self.gflat.numel() * 4
        starts = sorted((self.goffs[n], n) for n in self.grad_names)
        keys = [s for s, _ in starts]
        last = {}

        def tensors(obj):
            if isinstance(obj, torch.Tensor):
                yield obj
            elif isinstance(obj, (tuple, list)):
                for o in obj:
                    yield from tensors(o)
            elif isinstance(obj, dict):
                for o in obj.values():
                    yield from tensors(o)

        for i, op in enumerate(pl.bwd):
            if isinstance(op, str):
                continue
            for t in tensors(op.keep):
                p = t.data_ptr()
                if lo <= p < hi:
                    off = (p - lo) // 4
                    end = off + (t.numel() if t.is_contiguous() else 1)
                    j = bisect.bisect_right(keys, off) - 1
                    while j < len(starts) and starts[j][0] < end:
                        last[starts[j][1]] = i
                        j += 1
        return last

    # ------------------------------------------------------------------ execution
    def forward(self, B, inputs, noise, forced=None):
        """inputs: ids [B,T] int64, image [B,3,R,R] f32, seg [B,L] int64 (device tensors)."""
        if self.params_moved():
            raise L.SegclipB200Error("parameter storage moved after the engine was built; rebuild the engine")
        pl = self.plan(B)
        b = pl.bufs
        b["in.ids"].copy_(inputs["ids"], non_blocking=True)
        img = inputs["image"]
        self._img_event = None
        if img.device.type == "cpu":
            # host image (the reference's data contract): copy on a side stream so it overlaps the text tower; the
            # vision stem waits on the event ("wait_image" marker in the forward tape)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.dev)
            cur = torch.cuda.current_stream(self.dev)
            self._copy_stream.wait_stream(cur)           # previous step's readers of in.image are done
            with torch.cuda.stream(self._copy_stream):
                if img.dtype == torch.uint8:
                    self._upload_u8(img, b["in.image"], inputs["norm"])
                elif img.dtype == torch.float32 and img.dim() == 5:
                    b["in.image"].copy_(img[:, 0], non_blocking=True)
                else:
                    b["in.image"].copy_(img.reshape(b["in.image"].shape).to(torch.float32), non_blocking=True)
                self._img_event = self._copy_stream.record_event()
        elif img.dtype == torch.uint8:
            self._upload_u8(img, b["in.image"], inputs["norm"])
        else:
            b["in.image"].copy_(img.reshape(b["in.image"].shape), non_blocking=True)
        if inputs.get("seg") is not None:
            b["in.seg"].copy_(inputs["seg"].reshape(B, -1), non_blocking=True)
        b["in.u1"].copy_(noise["u1"], non_blocking=True)
        if self.use_mae:
            b["in.u2"].copy_(noise["u2"], non_blocking=True)
            b["in.u3"].copy_(noise["u3"], non_blocking=True)
        use_forced = forced is not None
        if use_forced:
            b["in.forced_main"].copy_(forced["main"].to(torch.int32))
            if self.use_mae:
                b["in.forced_mae"].copy_(forced["mae"].to(torch.int32))
        torch._foreach_zero_(pl.zero)
        self.gflat.zero_()
        st = L.stream()
        if self.cast_op is not None:
            self.cast_op(st)
        for op in self.prep_ops:
            op(st)
        self._refresh_conv_pad()
        for op in pl.fwd:
            if isinstance(op, str):
                if op == "force_pool":
                    if use_forced and forced.get("pool") is not None:
                        b["v.parg"].copy_(forced["pool"].to(torch.int32))
                elif op == "wait_image":
                    if self._img_event is not None:
                        torch.cuda.current_stream(self.dev).wait_event(self._img_event)
                else:
                    self._collective(op, pl)
            elif isinstance(op, tuple):
                op[1 if use_forced else 0](st)
            else:
                op(st)
        return pl.loss

    def _upload_u8(self, img, dst, norm):
        """uint8 pixels -> device (1 B/pixel over PCIe) -> normalised fp32 by the native kernel, on the current stream."""
        import ctypes
        if getattr(self, "_u8_stage", None) is None or self._u8_stage.numel() != dst.numel():
            self._u8_stage = torch.empty(dst.numel(), device=self.dev, dtype=torch.uint8)
        self._u8_stage.copy_(img.reshape(-1), non_blocking=True)
        mean = (ctypes.c_float * 3)(*norm[0])
        std = (ctypes.c_float * 3)(*norm[1])
        L.check(L.lib().sc_u8_normalize(self._u8_stage.data_ptr(), dst.data_ptr(), dst.numel(), dst.shape[-1] * dst.shape[-2],
                                        mean, std, L.stream()), "sc_u8_normalize")

    def infer(self, B, ids=None, image=None, norm=None):
        """Inference forward: no Gumbel noise, plain arg-max assignment (module_seg_vit.py:230-231),
        no losses.  Runs the text and/or visual tower and the hidden-state tails; results are read from plan buffers."""
        if self.params_moved():
            raise L.SegclipB200Error("parameter storage moved after the engine was built; rebuild the engine")
        pl = self.eval_plan(B)
        b = pl.bufs
        st = L.stream()
        if ids is not None:
            b["in.ids"].copy_(ids, non_blocking=True)
        if image is not None:
            if image.dtype == torch.uint8:
                self._upload_u8(image, b["in.image"], norm)
            else:
                b["in.image"].copy_(image.reshape(b["in.image"].shape).to(torch.float32), non_blocking=True)
        self._img_event = None
        torch._foreach_zero_(pl.zero)
        if self.cast_op is not None:
            self.cast_op(st)
        for op in self.prep_ops:
            op(st)
        self._refresh_conv_pad()
        # forward tape order: text tower, "wait_image", visual tower (+ head), "gather_embeddings", loss heads, MAE pass
        section = "text"
        for op in pl.fwd:
            if isinstance(op, str):
                if op == "wait_image":
                    section = "vision"
                elif op == "gather_embeddings":
                    break                      # towers done: loss heads / MAE pass are training-only
                continue
            if (section == "text" and ids is None) or (section == "vision" and image is None):
                continue
            (op[2] if isinstance(op, tuple) else op)(st)
        for which, given in (("text", ids), ("vision", image)):
            if given is not None:
                for op in pl.eval_tail[which]:
                    op(st)
        return b

    def backward(self, B):
        pl = self.plan(B)
        st = L.stream()
        works = []
        nv = self.nvls
        cur = torch.cuda.current_stream(self.dev) if pl.bucket_after else None

        def reduce_bucket(k):
            s, e_, _ = self.buckets[k]
            if nv is not None:         # mean over ranks by the NVLS kernel on its own stream, behind everything issued so far
                nv.all_reduce(s, e_, cur)
            else:
                import torch.distributed as dist
                works.append(dist.all_reduce(self.gflat[s:e_], group=self.sync_group, async_op=True))

        if pl.bucket_after:
            for k in pl.bucket_after.get(-1, []):          # parameters nobody writes (stay zero)
                reduce_bucket(k)
        for i, op in enumerate(pl.bwd):
            if isinstance(op, str):
                self._collective(op, pl)
            else:
                op(st)
            if pl.bucket_after and i in pl.bucket_after:
                for k in pl.bucket_after[i]:
                    reduce_bucket(k)
        for w in works:
            w.wait()               # stream-level wait, the host does not block
        if nv is not None and pl.bucket_after:
            if nv.profile is not None:                 # tools/sync_timeline.py: how long the step waits for the reductions
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur)
                nv.join(cur)
                e1.record(cur)
                nv.profile.append(("join", 0, (e0, e1)))
            else:
                nv.join(cur)
        return self.gflat

    def profile_gemm(self, B):
        """One extra instrumented step: CUDA events around every tensor-core GEMM launch (on the launching
        stream).  Returns total algorithmic FLOPs / total device time of the dominant kernel."""
        pl = self.plan(B)
        st = L.stream()
        recs = []
        dense = (self.kdense, self.vdense, self.dkdense, self.dvdense)

        def run(op):
            if isinstance(op, str):      # exchanges are skipped: this instrumented pass may run on one rank only
                return
            if isinstance(op, tuple):
                op = op[0]
            d = op.keep[0] if op.name == "sc_gemm" else None
            if d is not None and d.in_dtype == L.BF16:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                op(st)
                e1.record()
                fl = 2.0 * d.M * d.N * d.K
                if any(t is x for t in op.keep[1:4] for x in dense):
                    fl /= self.Hv          # grouped 1x1 conv run as a dense block-diagonal GEMM: algorithmic = 1/groups
                recs.append((e0, e1, fl))
            else:
                op(st)

        torch._foreach_zero_(pl.zero)
        self.gflat.zero_()
        for op in pl.fwd:
            run(op)
        for op in pl.bwd:
            run(op)
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b, _ in recs)
        fl = sum(f for _, _, f in recs)
        return dict(ms=ms, launches=len(recs), tflops=fl / ms / 1e9 if ms > 0 else 0.0, flops=fl)

    def _collective(self, what, pl):
        if self.world == 1:
            return
        if self.gather is None:
            raise L.SegclipB200Error("world_size > 1 but no embedding exchange is attached to the engine")
        if what == "gather_embeddings":
            self.gather.gather_embeddings(pl.slot)
        elif what == "gather_lse":
            self.gather.gather_lse(pl.slot)
        elif what == "release_exchange":
            self.gather.release(pl.slot)
        else:
            raise L.SegclipB200Error("unknown collective step %r" % what)
