"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (all launches in the file)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
n = 0
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[r[ui]]
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
    n += 1
print(f"{sys.argv[2] if len(sys.argv) > 2 else 'launch list'}: {n} launches, {tot / 1e3:.2f} ms (ncu per-launch durations, serialised)")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{c:4d}  {k}")
