"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
traffic_out = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--traffic" else None
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
tot = 0.0
dram = collections.defaultdict(lambda: [0, 0.0])        # kernel -> [launches, bytes]
for r in rows[1:]:
    if r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iu], 1.0)
        nm = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", ""))
        dram[nm][1] += float(r[iv].replace(",", "")) * mult
        if r[im] == "dram__bytes_read.sum":
            dram[nm][0] += 1
    if r[im] != "gpu__time_duration.sum":
        continue
    t = float(r[iv].replace(",", ""))
    t = t / 1e3 if r[iu] in ("ns", "nsecond") else t          # -> us
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    name = re.sub(r"<.*", "", name)
    c, s = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, s + t)
    tot += t
print("launches %d   total %.3f ms (cold-cache, serialised: compare SHARES)" % (sum(c for c, _ in agg.values()), tot / 1e3))
for name, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s x%-4d %9.3f ms  %5.1f%%" % (name, c, s / 1e3, 100 * s / tot))

if dram:
    print("--- DRAM traffic per launch (read + write)")
    for name, (c, b) in sorted(dram.items(), key=lambda kv: -kv[1][1])[:8]:
        print("%-34s x%-4d %9.1f MB/launch   %8.2f GB total" % (name, c, b / max(c, 1) / 1e6, b / 1e9))
    if traffic_out and "gemm_tc2_kernel" in dram:
        import json
        c, b = dram["gemm_tc2_kernel"]
        json.dump({"kernel": "gemm_tc2_kernel", "launches": c, "dram_bytes_per_launch": b / c,
                   "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one step (batch 256, contrastive)"},
                  open(traffic_out, "w"))
