"""Summarise the last N launches of an ncu launch list (csv with gpu__time_duration.sum [+ dram__bytes_read.sum]).
usage: python tools/launch_tail.py launches.csv [N]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 230
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ik, im, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
L = {}
for r in data:
    if len(r) <= iv:
        continue
    d = L.setdefault(r[0], {"k": r[ik], "g": r[ig]})
    d[r[im]] = float(r[iv].replace(",", ""))
ids = sorted(L, key=int)[-n_last:]
agg, total = OrderedDict(), 0.0
for i in ids:
    d = L[i]
    key = (d["k"].split("(")[0][-56:], d["g"])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0); a[2] += d.get("dram__bytes_read.sum", 0)
    total += d.get("gpu__time_duration.sum", 0)
print(f"{len(ids)} launches, {total / 1e3:.1f} us of kernel time")
for k, a in agg.items():
    print(f"{k[0]:58s} {k[1]:14s} n={a[0]:3d} avg_us={a[1] / a[0] / 1e3:7.2f} sum_us={a[1] / 1e3:8.1f} dram_MB={a[2] / a[0] / 1e6:8.2f}")
