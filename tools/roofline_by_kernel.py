"""Per-kernel roofline table from a tools/profile_step.py breakdown (CUDA-event times of every native call of one step).

    python tools/roofline_by_kernel.py profiles/r1_step_breakdown_cuda_events.txt > profiles/r1_roofline_by_kernel.txt

GEMM: algorithmic 2*M*N*K; attention: 4*B*H*Lq*Lk*hd forward, 2.5x that backward (5 products instead of 2), dense (causal
not halved, SURVEY 8(d) convention); LayerNorm: algorithmic bytes/row = D*(4+2) forward, D*(2+4+2+2) backward
(bf16 gradient stream).  Peaks: MEASURED_PEAKS.json (sustained bf16 cuBLAS, HBM copy)."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
TF, GB = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
rows = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+(fwd|bwd) x(\d+)\s+([\d.]+) ms\s+[\d.]+%\s+(.*)", line)
    if not m:
        continue
    ph, cnt, ms, desc = m.group(1), int(m.group(2)), float(m.group(3)), m.group(4).strip()
    kv = {k: int(v) for k, v in re.findall(r"(\w+)=(\d+)", desc)}
    if desc.startswith("gemm_tc"):
        fl = 2.0 * kv["M"] * kv["N"] * kv["K"] * cnt
        rows.append((ms, "%s %s" % (ph, desc), "tensor", fl / ms / 1e9, TF, "TFLOP/s"))
    elif desc.startswith("sc_attention"):
        fl = 4.0 * kv["B"] * kv["H"] * kv["Lq"] * kv["Lk"] * kv["hd"] * cnt * (2.5 if "bwd" in desc else 1.0)
        rows.append((ms, "%s %s" % (ph, desc), "tensor", fl / ms / 1e9, TF, "TFLOP/s"))
    elif desc.startswith("sc_layernorm"):
        by = kv["rows"] * kv["D"] * (10.0 if "bwd" in desc else 6.0) * cnt      # bwd: dy bf16 + x fp32 + bf16 gradient stream read-modify-write
        rows.append((ms, "%s %s" % (ph, desc), "hbm", by / ms / 1e6, GB, "GB/s"))
tot = sum(r[0] for r in rows)
print("# per-kernel roofline (one step, batch 256, ViT-B/16 contrastive); peaks: %.1f TFLOP/s sustained bf16, %.0f GB/s HBM" % (TF, GB))
print("# %-78s %9s %10s %7s" % ("kernel (calls per step)", "ms", "achieved", "frac"))
for ms, name, bound, ach, peak, unit in sorted(rows, key=lambda r: -r[0]):
    print("  %-78s %9.3f %8.1f %-8s %5.1f%%" % (name[:78], ms, ach, unit, 100.0 * ach / peak))
g = [r for r in rows if "gemm_tc" in r[1]]
print("# GEMM total %.2f ms, %.1f TFLOP/s = %.1f%% of peak" % (sum(r[0] for r in g), sum(r[3] * r[0] for r in g) / sum(r[0] for r in g),
                                                             100 * sum(r[3] * r[0] for r in g) / sum(r[0] for r in g) / TF))
