"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[1:]:
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
