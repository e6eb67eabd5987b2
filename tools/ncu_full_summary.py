#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` (a `--set full` capture) per kernel and refresh
profiles/traffic.json.  Usage: ncu_full_summary.py raw.csv frames_per_launch [traffic.json] [source-note] [HxW]"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
frames = float(sys.argv[2])
col = {h: i for i, h in reversed(list(enumerate(hdr)))}


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return v * scale


M = dict(t="gpu__time_duration.sum", inst="smsp__inst_executed.sum",
         issue="smsp__issue_active.avg.pct_of_peak_sustained_active",
         alu="sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         fma="sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         lsu="sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         xu="sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         warps="sm__warps_active.avg.pct_of_peak_sustained_active",
         rd="dram__bytes_read.sum", wr="dram__bytes_write.sum")
agg = collections.OrderedDict()
for r in data:
    k = r[col["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].replace("void ", "")
    a = agg.setdefault(k, collections.defaultdict(float))
    a["n"] += 1
    t = val(r, M["t"])
    for key in ("issue", "alu", "fma", "lsu", "xu", "warps"):
        a[key] += val(r, M[key]) * t          # time-weighted
    a["t"] += t
    a["inst"] += val(r, M["inst"])
    a["dram"] += val(r, M["rd"]) + val(r, M["wr"])
print("# columns: time us (sum over launches) | warp-instr per frame | issue-active % | ALU | FMA | LSU | XU pipe % | warps active % | DRAM read+write MB per frame")
traffic = {}
for k, a in agg.items():
    t = a["t"]
    print(f"{k:16s} launches={int(a['n']):d} time={t:8.1f} us  instr/frame={a['inst'] / frames / 1e6:6.3f} M  issue={a['issue'] / t:5.1f}%  "
          f"alu={a['alu'] / t:5.1f}%  fma={a['fma'] / t:5.1f}%  lsu={a['lsu'] / t:5.1f}%  xu={a['xu'] / t:5.1f}%  warps={a['warps'] / t:5.1f}%  "
          f"dram={a['dram'] / frames / 1e6:6.3f} MB/frame")
    traffic[k] = dict(dram_bytes_per_frame=a["dram"] / frames, alu_pipe_pct=a["alu"] / t, warp_instr_per_frame=a["inst"] / frames)
if len(sys.argv) > 3:
    # MERGE into profiles/traffic.json (keyed by frame size, then by bench stage name): entries of kernels that are not
    # in this capture stay as they are
    name = {"k_fast_cells": "fast", "k_blur": "blur", "k_describe": "describe", "k_octree<1>": "octree",
            "k_octree<0>": "octree", "k_layout": "layout"}
    size = sys.argv[5] if len(sys.argv) > 5 else "480x752"
    try:
        allt = json.load(open(sys.argv[3]))
    except Exception:
        allt = {}
    out = allt.setdefault(size, {})
    pyr = None
    for k, v in traffic.items():
        v["source"] = sys.argv[4] if len(sys.argv) > 4 else sys.argv[1]
        if k.startswith("k_level0") or k.startswith("k_resize") or k.startswith("k_pyr"):
            if pyr is None:
                pyr = dict(dram_bytes_per_frame=0.0, alu_pipe_pct=0.0, warp_instr_per_frame=0.0, source=v["source"])
            pyr["dram_bytes_per_frame"] += v["dram_bytes_per_frame"]
            pyr["warp_instr_per_frame"] += v["warp_instr_per_frame"]
            pyr["alu_pipe_pct"] = max(pyr["alu_pipe_pct"], v["alu_pipe_pct"])
        elif k in name:
            out[name[k]] = v
        else:
            out[k] = v
    if pyr is not None:
        out["pyramid"] = pyr
    json.dump(allt, open(sys.argv[3], "w"), indent=1, sort_keys=True)
