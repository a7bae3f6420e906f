#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
            ki, vi, ui = r.index("Kernel Name"), r.index("Metric Value"), r.index("Metric Unit")
        continue
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    nm = r[ki].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")[:64]
    a = agg[nm]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
tot = sum(v[1] for v in agg.values())
print("%-66s %5s %10s %6s %9s" % ("kernel", "n", "total us", "share", "max us"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s %5d %10.1f %5.1f%% %9.1f" % (k, v[0], v[1], 100 * v[1] / tot, v[2]))
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
