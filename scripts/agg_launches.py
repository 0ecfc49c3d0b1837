"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of scripts/one_forward.py by kernel
(second forward only) -> table printed to stdout."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
recs = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[r[ui]]
    recs.append((r[ki], v))
idx = [i for i, (k, _) in enumerate(recs) if "pack_input" in k]
fw = recs[idx[-1] - 2:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in fw:
    name = re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in fw)
print(f"one UNet forward, batch 8, 436x1024: {len(fw)} launches, {tot / 1e3:.2f} ms (ncu per-launch durations, serialised)")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{n:3d}  {k}")
