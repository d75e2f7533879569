"""Host-side guards that need no GPU: the product path never touches oracle/, and the product-side geometry helpers agree
with the oracle's."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(tree):
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name, node
        elif isinstance(node, ast.ImportFrom) and node.module:
            yield node.module, node


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "segclip_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            tree = ast.parse(open(os.path.join(pkg, fn)).read())
            bad = [m for m, _ in _imports(tree) if m == "oracle" or m.startswith("oracle.")]
            assert not bad, (fn, bad)


def test_bench_uses_the_oracle_only_in_the_cpu_legs():
    """Only `run_reference` (--impl reference) and `cpu_baseline` may import oracle/ (the checker timed beside the GPU)."""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"run_reference", "cpu_baseline", "_cpu_arm"}       # _cpu_arm: the shared body of the two CPU legs
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        uses = [m for m, _ in _imports(fn) if m == "oracle" or m.startswith("oracle.")]
        assert not uses or fn.name in allowed, (fn.name, uses)
        calls = [c for c in ast.walk(fn) if isinstance(c, ast.Call) and getattr(c.func, "id", None) == "_cpu_arm"]
        assert not calls or fn.name in ("run_reference", "cpu_baseline"), fn.name
    top = [m for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom)) for m, _ in _imports(n)]
    assert not [m for m in top if m.startswith("oracle")]


def test_geometry_presets_match_the_oracle():
    from oracle import segclip_oracle as so
    from oracle.ref_harness import fake_clip_state_dict
    from segclip_b200 import config
    assert config.vit_b16() == so.vit_b16_config()
    assert config.vit_b16(use_mae=True, use_kl=True) == so.vit_b16_config(use_mae=True, use_kl=True)
    for cfg in (config.vit_b16(), config.vit_l14(use_mae=True)):
        a, b = config.shape_state_dict(cfg), fake_clip_state_dict(cfg)
        assert a.keys() == b.keys()
        assert all(a[k].shape == b[k].shape for k in a)
