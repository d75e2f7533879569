"""The committed measurement evidence stays consistent with the tools that made it (no GPU needed): the ncu launch list of
one step re-summarises to the committed per-kernel shares, and the bench lines under profiles/ carry the contract's keys."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_launch_list_resummarises_to_the_committed_shares():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"),
                          os.path.join(PROF, "r2_launches_b256_contrastive.csv")], capture_output=True, text=True, check=True).stdout
    committed = open(os.path.join(PROF, "r2_launch_summary.txt")).read()
    assert out.strip().splitlines()[:6] == committed.strip().splitlines()[:6]
    top = out.splitlines()[1].split()
    assert top[0] == "gemm_tc2_kernel" and 55.0 < float(top[-1].rstrip("%")) < 75.0, top


def test_bench_lines_carry_the_contract_keys():
    files = sorted(glob.glob(os.path.join(PROF, "r2_bench_*.json")))
    assert len(files) >= 10, files
    for f in files:
        line = [l for l in open(f).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                  "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (os.path.basename(f), k)
        assert d["unit"] == "pairs/s" and d["dtype"] == "bf16" and d["scaling"] == "weak"
        assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"] + 1.0, os.path.basename(f)
        r = d["roofline"]
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]), os.path.basename(f)
