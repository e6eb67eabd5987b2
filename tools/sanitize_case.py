"""Small end-to-end case for compute-sanitizer (memcheck): one odd-sized frame, a stereo pair, the matchers."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")]
import numpy as np, synth, orbfe
ex = orbfe.ORBextractor(500)
for shape in [(241, 323), (480, 752)]:
    mono, k, d = ex(synth.synth_frame(shape[0], shape[1], 3), None, (0, 1000))
    print(shape, mono, len(k))
frames = np.stack([synth.synth_frame(240, 320, i) for i in range(5)])
ex.set_max_bytes(64 << 20)
n, mono, kps, desc = ex.extract_batch(frames, (0, 0))
print("batch", n)
l, r = synth.stereo_pair(240, 320, 1)
gl, gr = orbfe.ORBextractor(600), orbfe.ORBextractor(600)
_, kl, dl = gl(l, None, (0, 0)); _, kr, dr = gr(r, None, (0, 0))
ur, dp = orbfe.ORBmatcher.ComputeStereoMatches(gl, gr, kl, dl, kr, dr, 40.0, 0.1)
print("stereo", int((ur > 0).sum()))
m = orbfe.ORBmatcher()
print("knn", m.knn2(dl, dr)[2][:5])
d5 = synth.map_vs_frame(5000, 600, 1, w=320, h=240)
rng = np.random.default_rng(0)
pts = dict(u=d5["u"], v=d5["v"], ur=d5["u"], radius=np.full(5000, 12, np.float32), min_level=np.zeros(5000, np.int32),
           max_level=np.full(5000, -1, np.int32), angle=np.zeros(5000, np.float32), valid=np.ones(5000, np.uint8),
           blocks=np.ones(5000, np.uint8), desc=d5["mdesc"])
F = orbfe.FrameData(d5["keys"], d5["fdesc"], d5["bounds"], None)
print("search", m.SearchByProjection(F, pts, np.zeros(600, np.uint8), np.full(600, -1, np.int32))[0])
