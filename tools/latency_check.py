"""Per-call latency of the projection searches (host arrays in and out), as bench.py's call_latency_ms measures them."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam3_byzyh_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import orbfe
import synth


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


m = orbfe.ORBmatcher(0.8, True)
for n_pts in (1000, 3000, 20000):
    d = synth.map_vs_frame(n_pts, 1200, 1, w=752, h=480)
    pts = dict(u=d["u"], v=d["v"], ur=d["u"], radius=np.full(n_pts, 10, np.float32), min_level=np.zeros(n_pts, np.int32),
               max_level=np.full(n_pts, -1, np.int32), angle=np.zeros(n_pts, np.float32), valid=np.ones(n_pts, np.uint8),
               blocks=np.ones(n_pts, np.uint8), desc=d["mdesc"])
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], None)
    cl, asg = np.zeros(1200, np.uint8), np.full(1200, -1, np.int32)
    print(n_pts, "SearchByProjection %.3f ms" % timeit(lambda: m.SearchByProjection(F, pts, cl, asg)),
          "LastFrame %.3f ms" % timeit(lambda: m.SearchByProjectionLastFrame(F, pts, cl, asg)),
          "Fuse %.3f ms" % timeit(lambda: m.FuseSearch(F, pts)))
