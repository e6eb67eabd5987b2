"""Where one ORBextractor::operator() call (one 752x480 frame, host arrays in and out) spends its time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam3_byzyh_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orbfe
import synth

img = synth.synth_frame(480, 752, 3)
ex = orbfe.ORBextractor(1000)
for _ in range(10):
    ex(img, None, (0, 1000))
t0 = time.perf_counter()
for _ in range(100):
    ex(img, None, (0, 1000))
print("wall per call %.3f ms" % ((time.perf_counter() - t0) / 100 * 1e3))
ex.set_profiling(True)
acc = {}
for _ in range(50):
    ex(img, None, (0, 1000))
    for k, v in ex.stage_ms().items():
        acc[k] = acc.get(k, 0.0) + v / 50
print({k: round(v, 4) for k, v in acc.items()}, "sum %.3f" % sum(acc.values()))
