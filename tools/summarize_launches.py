#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total,
share of the step, average).  Usage: summarize_launches.py launches.csv [skip_first_n [count]]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
agg = collections.OrderedDict()
for r in rows[1 + skip:1 + skip + count]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki].replace("<unnamed>::", "").split("(")[0].replace("void ", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':28s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
for k, a in agg.items():
    print(f"{k:28s} {a[0]:8d} {a[1] / 1e6:10.3f} {a[1] / tot * 100:6.1f}% {a[1] / a[0] / 1e3:10.1f}")
print(f"{'total':28s} {sum(a[0] for a in agg.values()):8d} {tot / 1e6:10.3f}")
