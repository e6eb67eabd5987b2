import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "orb-slam3_byzyh_b200")]
import numpy as np, torch, orbfe
from orbfe import _lib
L = orbfe.lib()
nq, nt = 2000, int(os.environ.get("NT", "1000000"))
rng = np.random.default_rng(1)
q = torch.from_numpy(rng.integers(0, 256, (nq, 32), dtype=np.uint8)).cuda()
t = torch.from_numpy(rng.integers(0, 256, (nt, 32), dtype=np.uint8)).cuda()
idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(L.orbfe_knn2_device(_lib.ptr(q), nq, _lib.ptr(t), nt, 0, _lib.ptr(idx), _lib.ptr(dist), st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.check(L.orbfe_knn2_device(_lib.ptr(q), nq, _lib.ptr(t), nt, 0, _lib.ptr(idx), _lib.ptr(dist), st))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"knn2 2000 x {nt}: {ms:.3f} ms  {nq * nt / ms / 1e6:.1f} G pairs/s")
