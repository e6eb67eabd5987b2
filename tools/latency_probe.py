"""Per-call latencies of the drop-in entry points at SLAM-frame sizes (host pointers in/out)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")]
import numpy as np, synth, orbfe


def timeit(fn, n=30):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3


left, right = synth.stereo_pair(480, 752, 3)
gl, gr = orbfe.ORBextractor(1200), orbfe.ORBextractor(1200)
_, kl, dl = gl(left, None, (0, 0)); _, kr, dr = gr(right, None, (0, 0))
print(f"extract 752x480 nF=1200 (one call)      : {timeit(lambda: gl(left, None, (0, 0))):.3f} ms")
print(f"ComputeStereoMatches 1200+1200 kps      : {timeit(lambda: orbfe.ORBmatcher.ComputeStereoMatches(gl, gr, kl, dl, kr, dr, 47.9, 0.11)):.3f} ms")
m = orbfe.ORBmatcher(0.8, True)
print(f"kNN-2 + ratio 1500 x 1500               : {timeit(lambda: m.knn2(dl, dr)):.3f} ms")
d = synth.map_vs_frame(3000, 1200, 1, w=752, h=480)
rng = np.random.default_rng(0)
pts = dict(u=d["u"], v=d["v"], ur=d["u"], radius=np.full(3000, 10, np.float32), min_level=np.zeros(3000, np.int32),
           max_level=np.full(3000, -1, np.int32), angle=np.zeros(3000, np.float32), valid=np.ones(3000, np.uint8),
           blocks=np.ones(3000, np.uint8), desc=d["mdesc"])
F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], None)
cl, asg = np.zeros(1200, np.uint8), np.full(1200, -1, np.int32)
print(f"SearchByProjection 3000 pts x 1200 kps  : {timeit(lambda: m.SearchByProjection(F, pts, cl, asg)):.3f} ms")
print(f"SearchByProjection (last frame, rot.)   : {timeit(lambda: m.SearchByProjectionLastFrame(F, pts, cl, asg)):.3f} ms")
