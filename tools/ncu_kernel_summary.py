#!/usr/bin/env python
"""Per-kernel averages of a `ncu --set full` capture (`ncu -i X.ncu-rep --page raw --csv`): time per launch, warp
instructions per launch, issue / pipe utilisation, warps active, DRAM bytes per launch.  Usage: ncu_kernel_summary.py raw.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in reversed(list(enumerate(hdr)))}


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(units[col[name]], 1.0)


PCT = dict(issue="smsp__issue_active.avg.pct_of_peak_sustained_active",
           alu="sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           fma="sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           lsu="sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           xu="sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           warps="sm__warps_active.avg.pct_of_peak_sustained_active")
agg = collections.OrderedDict()
for r in data:
    k = r[col["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].replace("void ", "")
    a = agg.setdefault(k, collections.defaultdict(float))
    t = val(r, "gpu__time_duration.sum")
    a["n"] += 1
    a["t"] += t
    a["inst"] += val(r, "smsp__inst_executed.sum")
    a["dram"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    a["grid"] = r[col["Grid Size"]]
    a["block"] = r[col["Block Size"]]
    for key, name in PCT.items():
        a[key] += val(r, name) * t
print("# per launch (averages over the launches of the capture): time us | warp-instr M | issue-active % | ALU | FMA | LSU | XU pipe % | warps active % | DRAM read+write MB")
for k, a in agg.items():
    n, t = a["n"], a["t"]
    print(f"{k:24s} launches={int(n):2d} grid={a['grid']:>14s} block={a['block']:>12s} time={t / n:8.1f} us  instr={a['inst'] / n / 1e6:8.3f} M  "
          + "  ".join(f"{key}={a[key] / t:5.1f}%" for key in PCT) + f"  dram={a['dram'] / n / 1e6:8.3f} MB")
