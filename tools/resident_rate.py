"""Device-resident extraction rate (ms per batch) under a list of environment settings, one process.
usage: resident_rate.py "K1=V1,K2=V2" "K1=V3" ...   (knobs the library reads per call)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")]
import torch, orbfe, bench
B = int(os.environ.get("RR_FRAMES", "1024"))
frames = torch.from_numpy(bench.make_frames(B)).cuda()
ex = orbfe.ORBextractor(1000); ex.set_max_bytes(64 << 30)
cap = ex.capacity
o = [torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda"), torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"),
     torch.empty(B, dtype=torch.int32, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda")]
st = torch.cuda.Stream()
def run(steps=10):
    with torch.cuda.stream(st):
        for _ in range(3): ex.extract_batch_device(frames, (0, 1000), o[0], o[1], o[2], o[3], st)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps): ex.extract_batch_device(frames, (0, 1000), o[0], o[1], o[2], o[3], st)
        e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / steps
base = dict(os.environ)
for spec in sys.argv[1:] or [""]:
    os.environ.clear(); os.environ.update(base)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("="); os.environ[k] = v
    ms = run()
    print(f"{spec or '(default)':40s} {ms:.3f} ms per {B} frames -> {B / ms * 1e3:.0f} frames/s  checksum {int(o[2].sum())}")
