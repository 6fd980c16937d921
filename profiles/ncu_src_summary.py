#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` per kernel: stall-reason totals and the hottest SASS lines.
usage: ncu -i rep --page source --csv | python profiles/ncu_src_summary.py [kernel-substring] [ntop]"""
import csv, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ""
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rows = list(csv.reader(sys.stdin))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr) - 2: body.append(rows[j])
            j += 1
        if pat in name:
            print("==", name[:120])
            si, ns = hdr.index("Source"), hdr.index("# Samples")
            cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            tot = {hdr[k]: sum(float(r[k] or 0) for r in body) for k in cols}
            T = sum(tot.values()) or 1
            print("  stalls:", ", ".join(f"{k[6:]} {100*v/T:.1f}%" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
            for r in sorted(body, key=lambda r: -float(r[ns] or 0))[:ntop]:
                print(f"  {int(float(r[ns] or 0)):6d}  {r[si].strip()[:100]}")
        i = j
    else:
        i += 1
