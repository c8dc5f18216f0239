#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line and by device function.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; ncu_by_line.py src.csv fp_device.cuh [units]"""
import bisect
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
srcfile = sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr = None
data = []
other = {}
cur_file = ""
import os
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); ism = hdr.index("# Samples")
        continue
    if hdr and r and r[0].isdigit() and len(r) > ie:
        try:
            rec = (int(r[0]), r[1], int(r[ie] or 0), int(r[ism] or 0))
        except ValueError:
            continue
        if os.path.basename(cur_file) == os.path.basename(srcfile):
            data.append(rec)
        else:
            o = other.setdefault(os.path.basename(cur_file) + ":" + r[1].strip()[:60], [0, 0]); o[0] += rec[2]; o[1] += rec[3]
tot = sum(d[2] for d in data) + sum(o[0] for o in other.values()); tots = sum(d[3] for d in data) + sum(o[1] for o in other.values())
src = open(srcfile).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    if l.startswith("__device__") or l.startswith("__global__"):
        name = re.findall(r"(\w+)\s*\(", l)
        funcs.append((i, name[0] if name else l[:30]))
starts = [f[0] for f in funcs]
agg = {}
for ln, s, n, sm in data:
    k = bisect.bisect_right(starts, ln) - 1
    name = funcs[k][1] if k >= 0 else "?"
    a = agg.setdefault(name, [0, 0]); a[0] += n; a[1] += sm
print(f"total warp-inst {tot}  samples {tots}  inst/unit {tot/units:.1f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} inst {v[0]/tot*100:5.1f}%  samples {v[1]/tots*100:5.1f}%  inst/unit {v[0]/units:8.1f}")
print("-- lines in other files (headers) --")
for k, v in sorted(other.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"{k:70s} inst {v[0]/tot*100:5.1f}%  samples {v[1]/tots*100:5.1f}%  inst/unit {v[0]/units:8.1f}")
print()
for ln, s, n, sm in sorted(data, key=lambda d: -d[3])[:45]:
    print(f"{ln:5d} {n/tot*100:5.1f}% inst {sm/tots*100:5.1f}% smp  {s.strip()[:120]}")
