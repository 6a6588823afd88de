#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: launch_summary.py launches.csv [skip_first_n_launches]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
tot = 0.0
for r in rows[1:]:
    if int(r[idi]) < skip:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].replace("<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(name, [0.0, 0]); a[0] += v; a[1] += 1; tot += v
print(f"total {tot:.3f} ms over {sum(a[1] for a in agg.values())} launches")
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:10.3f} ms {100*ms/tot:6.2f}% {n:5d}  {name[:110]}")
