"""Text summary of one `ncu --set full` report: the headline metrics plus the top stall sites of the SASS view.

    python tools/ncu_summary.py report.ncu-rep [title] > profiles/r1_ncu_<kernel>.txt
"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
WANT = [r"^gpu__time_duration\.sum$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
        r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
        r"^sm__inst_executed_pipe_tensor.*hmma\.avg\.pct", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^sm__issue_active\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__inst_executed_pipe_xu\.avg\.pct_of_peak_sustained_active$",
        r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__data_pipe_tc_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^lts__t_sector_hit_rate\.pct$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_.*)$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__cycles_elapsed\.avg\.per_second$", r"^smsp__inst_executed\.sum$", r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
print("# ncu --set full --clock-control none :: %s" % title)
print("# kernel: %s" % v[h.index("Kernel Name")])
for i, n in enumerate(h):
    if any(re.search(w, n) for w in WANT):
        print("%-92s %-18s %s" % (n, u[i], v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h, data = rows[1], rows[2:]
isrc, iall = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
stallcols = [i for i, nm in enumerate(h) if nm.startswith("stall_") and "Not Issued" not in nm]
tot = sum(int(r[iall] or 0) for r in data)
ops = {k: sum(1 for r in data if k in r[isrc]) for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA")}
print("\n# SASS: %d instructions; %s" % (len(data), ", ".join("%s x%d" % kv for kv in ops.items() if kv[1])))
print("# top stall sites (%d warp samples)" % tot)
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iall] or 0))[:14]):
    r = data[i]
    st = sorted(((int(r[c] or 0), h[c]) for c in stallcols), reverse=True)[0]
    print("%5.1f%%  %-78s %s" % (100.0 * int(r[iall] or 0) / max(tot, 1), r[isrc].strip()[:78], st[1]))
