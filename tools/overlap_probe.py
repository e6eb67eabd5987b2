"""How much do two concurrent half-batches (two extractor instances on two streams) gain over one
full batch?  Decides whether the library should split a device batch over two internal streams."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")]
import torch, numpy as np, orbfe
import bench
B = 512
frames = torch.from_numpy(bench.make_frames(B)).cuda()
def mk(n):
    ex = orbfe.ORBextractor(1000); ex.set_max_bytes(64 << 30)
    cap = ex.capacity
    return ex, [torch.empty((n, cap, 28), dtype=torch.uint8, device="cuda"), torch.empty((n, cap, 32), dtype=torch.uint8, device="cuda"),
                torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")]
def run(parts, steps=10):
    exs = [mk(len(p)) for p in parts]
    sts = [torch.cuda.Stream() for _ in parts]
    def step():
        for (ex, o), p, st in zip(exs, parts, sts):
            ex.extract_batch_device(p, (0, 1000), o[0], o[1], o[2], o[3], st)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in sts: st.wait_event(e0)
    for _ in range(steps): step()
    for st in sts: torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for k in (1, 2, 4):
    parts = list(frames.chunk(k))
    ms = run(parts)
    print(f"{k} concurrent stream(s) x {B // k} frames: {ms:.3f} ms per {B} frames -> {B / ms * 1e3:.0f} frames/s")
