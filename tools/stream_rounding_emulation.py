"""Design experiment (CPU, oracle only -- test infrastructure): what does a bf16 RESIDUAL STREAM cost in gradient parity?

Re-runs the fp32 oracle with every matrix-product operand rounded to bf16 (oracle/bf16_emulation.py, the irreducible floor
of the CUDA path) and additionally rounds (a) the forward residual stream after every residual add, (b) the gradient
arriving at every residual add, to bf16.  Result on ViT-B/16, B = 8 (profiles/r2_stream_rounding_emulation.txt): the
forward rounding triples the worst per-tensor gradient error (bias / LayerNorm gradients 0.10 -> 0.27-0.35 rel-L2), the
backward rounding leaves it unchanged -> the engine keeps the forward stream fp32 and stores the gradient stream in bf16.

    python tools/stream_rounding_emulation.py [batch]
"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import segclip_oracle as so
from oracle.bf16_emulation import BF16Operands, _RoundFwd, _RoundBwd
from segclip_b200.engine import FROZEN_STEM

class Mode(BF16Operands):
    ADDS = {torch.Tensor.add, torch.Tensor.__add__, torch.add}
    def __init__(self, fwd, bwd):
        super().__init__(); self.f, self.b = fwd, bwd; self.n = 0
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in self.ADDS and len(args) == 2 and all(isinstance(a, torch.Tensor) for a in args) and args[0].dim() == 3 \
                and args[0].shape == args[1].shape and args[0].shape[-1] in (512, 768) and args[0].shape[1] in (77, 196, 197):
            out = func(*args, **kwargs)
            self.n += 1
            if self.f: out = _RoundFwd.apply(out)
            if self.b and out.requires_grad: out = _RoundBwd.apply(out)
            return out
        return super().__torch_function__(func, types, args, kwargs)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = so.vit_b16_config()
params = so.init_params(cfg, seed=5)
batch, noise = so.make_batch(cfg, B, seed=6)
ref_loss, ref, info = so.loss_and_grads(params, batch, noise, cfg, "torch18_flat", frozen=FROZEN_STEM)
f = {"main": info["assign_main"], "pool": info["pool_arg"]}
def run(fwd, bwd):
    m = Mode(fwd, bwd)
    with m:
        l, g, _ = so.loss_and_grads(params, batch, noise, cfg, "torch18_flat", forced=f, frozen=FROZEN_STEM)
    errs = {k: float((g[k]-ref[k]).norm())/(float(ref[k].norm())+1e-12) for k in ref if k in g}
    v = sorted(errs.values())
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print("fwd=%d bwd=%d adds=%d loss_rel=%.2e median=%.4f max=%.4f" % (fwd, bwd, m.n, abs(float(l)-float(ref_loss))/abs(float(ref_loss)), v[len(v)//2], v[-1]), [(k.split('.')[-3:], round(e,3)) for k,e in worst])
    return errs
run(0,0); run(1,0); run(0,1); run(1,1)
