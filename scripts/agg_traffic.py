"""Aggregate an ncu launch list of scripts/one_forward.py captured with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file <csv> python scripts/one_forward.py
by kernel -> table on stdout; with --json <path> <workload> also records the DRAM traffic of the tensor-core conv launches of
the forward (bench.py's roofline.traffic reads it from profiles/r2_forward_traffic.json)."""
import collections
import csv
import json
import re
import sys

args = sys.argv[1:]
rows = list(csv.reader(open(args[0])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ii, ki, mi, vi, ui = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    if unit in ("ns", "us", "ms", "s"):
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[unit]
    elif unit in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    launch.setdefault(r[ii], {"name": r[ki]})[r[mi]] = v
recs = list(launch.values())
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in recs:
    name = re.sub(r"\(.*", "", r["name"]).replace("void ", "").replace("<unnamed>::", "")
    a = agg[name]
    a[0] += 1
    a[1] += r.get("gpu__time_duration.sum", 0.0)
    a[2] += r.get("dram__bytes_read.sum", 0.0)
    a[3] += r.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
rd, wr = sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
print(f"one UNet forward: {len(recs)} launches, {tot / 1e3:.2f} ms (ncu per-launch durations, serialised), DRAM read {rd / 1e9:.2f} GB "
      f"write {wr / 1e9:.2f} GB")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    bw = (a[2] + a[3]) / max(a[1], 1e-9) / 1e6
    print(f"{a[1]:10.1f} us {100 * a[1] / tot:5.1f}%  x{a[0]:3d}  read {a[2] / 1e6:9.1f} MB  write {a[3] / 1e6:9.1f} MB  {bw:5.2f} TB/s  {k}")
conv = [a for k, a in agg.items() if k.startswith(("conv_igemm", "conv3x3_strip"))]
cb = sum(a[2] + a[3] for a in conv)
ct = sum(a[1] for a in conv)
print(f"all conv kernels: {sum(a[0] for a in conv)} launches, {ct:.1f} us, DRAM traffic {cb / 1e9:.3f} GB per forward")
if "--json" in args:
    path, wl = args[args.index("--json") + 1], args[args.index("--json") + 2]
    try:
        d = json.load(open(path))
    except Exception:  # noqa: BLE001
        d = {}
    d[wl] = {"conv_dram_bytes": cb, "conv_us_under_ncu": ct, "forward_dram_bytes": rd + wr, "forward_us_under_ncu": tot,
             "launches": len(recs), "source": args[0].split("/")[-1] + " (ncu launch list of scripts/one_forward.py, one batch forward)"}
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
