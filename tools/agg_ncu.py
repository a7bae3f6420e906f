#!/usr/bin/env python
"""Aggregates an `ncu --csv` metric log of one step by kernel name: launches, total/avg us, DRAM GB/s and % of peak,
SM throughput %, achieved occupancy %, tensor-pipe %, L2 atomic sectors.  python tools/agg_ncu.py step.csv"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    key = r["ID"]
    d = per.setdefault(key, {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
    try:
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d[r["Metric Name"] + "#unit"] = r["Metric Unit"]
    except ValueError:
        pass


def us(d):
    v = d.get("gpu__time_duration.sum", 0.0)
    u = d.get("gpu__time_duration.sum#unit", "ns")
    return v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)


def nbytes(d, k):
    v = d.get(k, 0.0)
    u = d.get(k + "#unit", "byte")
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


agg = collections.OrderedDict()
for d in per.values():
    nm = d["name"].replace("(anonymous namespace)::", "").replace("void ", "")[:58]
    a = agg.setdefault(nm, {"n": 0, "us": 0.0, "bytes": 0.0, "smw": 0.0, "occw": 0.0, "tcw": 0.0, "atom": 0.0, "maxus": 0.0})
    t = us(d)
    a["n"] += 1
    a["us"] += t
    a["maxus"] = max(a["maxus"], t)
    a["bytes"] += nbytes(d, "dram__bytes_read.sum") + nbytes(d, "dram__bytes_write.sum")
    a["smw"] += t * d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
    a["occw"] += t * d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
    a["tcw"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    a["atom"] += d.get("lts__t_sectors_op_atom.sum", 0.0) + d.get("lts__t_sectors_op_red.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
print("%-58s %4s %9s %6s %7s %8s %5s %5s %5s %9s" % ("kernel", "n", "total us", "share", "max us", "DRAM GB/s", "SM%", "occ%", "TC%", "L2 atom"))
for nm, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    t = a["us"] or 1e-9
    print("%-58s %4d %9.1f %5.1f%% %7.1f %8.0f %5.1f %5.1f %5.1f %9.0f" % (
        nm, a["n"], a["us"], 100 * a["us"] / tot, a["maxus"], a["bytes"] / t / 1e3, a["smw"] / t, a["occw"] / t, a["tcw"] / t, a["atom"]))
print("total %.1f us over %d launches" % (tot, sum(a["n"] for a in agg.values())))
