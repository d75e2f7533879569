"""Helpers shared by the CPU (oracle) and GPU (CUDA path) golden-vector tests."""
import hashlib
import json
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        return json.load(f)


def checksum_vector(name, numel):
    seed = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
    return torch.randn(numel, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


def compare_grads(grads, golden, tol, skip=()):
    """golden: {name: [l2 norm, checksum]}.  Returns list of (name, norm_err, chk_err) failures.

    A gradient g with relative error eps has |<g-g*, r>| ~ eps*|g*| for a unit-variance random r,
    so both checks are scaled by the golden norm."""
    bad = []
    for name, (norm, chk) in golden.items():
        if name in skip:
            continue
        if name not in grads:
            bad.append((name, "missing", None))
            continue
        g = grads[name].detach().double().cpu().flatten()
        scale = max(norm, 1e-12)
        e_norm = abs(float(g.norm()) - norm) / scale
        e_chk = abs(float(g @ checksum_vector(name, g.numel())) - chk) / scale
        if e_norm > tol or e_chk > tol:
            bad.append((name, e_norm, e_chk))
    return bad
