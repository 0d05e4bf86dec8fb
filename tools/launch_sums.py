#!/usr/bin/env python
"""Sum of gpu__time_duration per kernel name from an ncu --csv launch list: python tools/launch_sums.py file.csv"""
import csv
import collections
import sys

tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
with open(sys.argv[1]) as f:
    rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = r[ki].split("(")[0]
    t = tot[name]
    t[0] += 1
    t[1] += v
    t[2] = max(t[2], v)
for name, (n, us, mx) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:60s} launches {n:6d}  total {us / 1e3:10.2f} ms  mean {us / n:9.2f} us  max {mx:9.2f} us")
