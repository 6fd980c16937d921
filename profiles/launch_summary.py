#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the LAST forward (`n` launches of our kernels).
usage: python profiles/launch_summary.py gpurun_out/launches.csv [launches_per_step]"""
import csv, re, sys
from collections import OrderedDict
path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 69
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    if name.startswith("void at::") or "at::native" in name or "nccl" in name.lower():
        continue
    short = re.sub(r"\(.*$", "", name).replace("void ", "")
    rows.append((short, r["Grid Size"], float(r["Metric Value"]) / 1e3))
rows = rows[-n:]
tot = sum(t for _, _, t in rows)
agg = OrderedDict()
for k, g, t in rows:
    a = agg.setdefault(k, [0.0, 0]); a[0] += t; a[1] += 1
print(f"# total {tot:.1f} us over {len(rows)} launches (cold-cache, serialised: compare SHARES)")
for k, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}% x{c:3d}  {k}")
print("\n# launch order")
for k, g, t in rows:
    print(f"{t:10.1f} us  {g:>16}  {k}")
