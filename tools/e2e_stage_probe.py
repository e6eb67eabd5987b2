#!/usr/bin/env python
"""tools/e2e_stage_probe.py -- do the kernels run slower while host<->device copies are in flight?  Per-stage CUDA-event
times (orbfe_set_profiling: one compute stream) of 1024-frame launches, (a) frames resident in HBM, nothing else running,
(b) inside the streaming host path (orbfe_extract_batch_submit/_wait, two batches in flight: the H2D of batch k+1 and
the D2H of batch k-1 run under the kernels of batch k).  Also samples SM clocks during (b)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")):
    sys.path.insert(0, p)
import bench
import orbfe
from orbfe import _lib

B, LAP = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, (0, 1000)
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else B
dev = torch.device("cuda:0")
frames = bench.make_frames(B, seed0=7)
pinned = torch.from_numpy(frames).pin_memory()
ex = orbfe.ORBextractor(1000, 1.2, 8, 20, 7, device=0)
cap = ex.capacity
ex.set_max_bytes(64 << 30)
d_img = pinned.to(dev)
d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.empty(B, dtype=torch.int32, device=dev)
d_mono = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(device=dev)
for _ in range(3):
    ex.extract_batch_device(d_img, LAP, d_kps, d_desc, d_n, d_mono, st)
st.synchronize()
ex.set_profiling(True)
for _ in range(6):
    ex.extract_batch_device(d_img, LAP, d_kps, d_desc, d_n, d_mono, st)
st.synchronize()
a = ex.stage_ms()
ex.set_profiling(False)
geo = ex.frame_geometry()
ex.set_max_bytes(int(geo["per_frame_bytes"]) * chunk)


def outs():
    return (torch.empty(B, dtype=torch.int32).pin_memory().numpy(), torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
            torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory().numpy().view(_lib.KP_DTYPE).reshape(B, cap),
            torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory().numpy())


o = (outs(), outs())
ins = (pinned, pinned.clone().pin_memory())


def stream_steps(k):
    ex.extract_batch_submit(ins[0], LAP, o[0])
    for i in range(1, k):
        ex.extract_batch_submit(ins[i & 1], LAP, o[i & 1])
        ex.extract_batch_wait()
    ex.extract_batch_wait()


stream_steps(3)
clocks = bench.ClockSampler(0)
clocks.start()
t0 = time.time()
while len(clocks.lines) < 4 and time.time() - t0 < 6:
    stream_steps(4)
ex.set_profiling(True)
t0 = time.perf_counter()
stream_steps(8)
dt = (time.perf_counter() - t0) / 8 * 1e3
b = ex.stage_ms()
ex.set_profiling(False)
clk = clocks.stop()
print(f"frames per step {B}, chunk {chunk}; streaming step {dt:.3f} ms (profiling on: one compute stream)")
print(f"{'stage':12s} {'resident ms':>12s} {'streaming ms':>13s}   (per step)")
scale = B / chunk
for k in a:
    print(f"{k:12s} {a[k]:12.3f} {b[k] * scale:13.3f}")
print(f"{'sum':12s} {sum(a.values()):12.3f} {sum(b.values()) * scale:13.3f}")
print("clocks during streaming:", clk)
