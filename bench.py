#!/usr/bin/env python
"""bench.py -- ORB extraction frames/s (752x480, 1000 kp, 8 levels) and Hamming matches/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the ORB front-end over a batch of synthetic 752x480 frames per GPU
(EuRoC mono settings: nFeatures 1000, scaleFactor 1.2, 8 levels, FAST 20/7, lapping {0,1000}).
  value : frames/s with the batch resident in HBM (orbfe_extract_batch_device on a CUDA stream,
          CUDA events on that stream, max over ranks).
  e2e   : frames/s through the host-pointer C ABI (pinned host frames in, keypoints + descriptors out; the
          H2D/D2H copies of every step inside the timed region), driven as a streaming caller does:
          orbfe_extract_batch_submit / _wait with two batches in flight.  e2e.sync_call_value is the
          same with one blocking orbfe_extract_batch call per step.
Frames shard over ranks with no collective (weak scaling).  The matching leg (C5: 2000 frame
descriptors vs a 1 M descriptor map, sharded over ranks, NCCL all-gather of the per-shard best
two + merge) is reported in the same JSON line under "matching".
--impl reference times the reference's own CPU ORBextractor (oracle/_ref, its ORBextractor.cc
compiled verbatim; falls back to the oracle port) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W, NF = 480, 752, 1000
LAP = (0, 1000)
METRIC = "ORB extract frames/s (752x480,1000kp,8lvl)"
# BASELINE.json configs that are throughput workloads: C1 (the one `metric` is quoted on; the default) and C4.
WORKLOADS = {
    "c1": dict(h=480, w=752, nf=1000, frames=1024, metric=METRIC,
               label="C1 752x480 nFeatures=1000 scaleFactor=1.2 nLevels=8 FAST 20/7 lapping {0,1000}"),
    "c4": dict(h=720, w=1280, nf=2000, frames=384, metric="ORB extract frames/s (1280x720,2000kp,8lvl)",
               label="C4 1280x720 nFeatures=2000 scaleFactor=1.2 nLevels=8 FAST 20/7 lapping {0,1000}"),
}


def set_workload(name):
    global H, W, NF, METRIC
    wl = WORKLOADS[name]
    H, W, NF, METRIC = wl["h"], wl["w"], wl["nf"], wl["metric"]
    return wl


def make_frames(n, seed0=7, nbase=12):
    """n distinct H x W frames: `nbase` synthetic scenes (tests/synth.py), the rest are
    translated / mirrored variants (different pixel content per frame, same statistics)."""
    import synth
    base = [synth.synth_frame(H, W, seed0 + i) for i in range(min(nbase, n))]
    rng = np.random.default_rng(seed0)
    out = np.empty((n, H, W), np.uint8)
    for i in range(n):
        f = base[i % len(base)]
        if i >= len(base):
            f = np.roll(f, (int(rng.integers(1, H)), int(rng.integers(1, W))), axis=(0, 1))
            if rng.uniform() < 0.5:
                f = f[:, ::-1]
        out[i] = f
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(frames_per_thread, threads, frames):
    """Times the reference CPU ORBextractor (one instance per thread, as Frame.cc:136-141 does for
    stereo) on `threads` host threads.  Returns (frames/s, kind)."""
    try:
        from oracle import ref as R
        if R.available():
            R.lib()
            mk, kind = (lambda: R.RefExtractor(NF)), "reference"
        else:
            raise RuntimeError
    except Exception:
        from oracle import oracle as O
        O.lib()
        mk, kind = (lambda: O.Extractor(NF)), "port"
    exs = [mk() for _ in range(threads)]

    def work(t):
        for i in range(frames_per_thread):
            exs[t](frames[(t * frames_per_thread + i) % len(frames)], LAP)   # ctypes releases the GIL
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    return threads * frames_per_thread / dt, kind


def cv2_primitive_ms(frame, reps=5):
    """The OpenCV primitives the reference links (resize, copyMakeBorder, FAST, GaussianBlur: ORBextractor.cc:1702,
    1712, 1135, 1632) through cv2's own SIMD code, ONE thread, whole-level calls: the 'what the reference actually links'
    figure of SURVEY 8(d).  Not the reference's control flow (it makes one cv::FAST call per 35-px cell, and the octree,
    orientation and descriptors are its own scalar code): a lower bound for the three stages it covers."""
    try:
        import cv2
    except Exception as e:
        return {"error": "cv2 not importable: " + str(e)[:80]}
    cv2.setNumThreads(1)
    inv = [1.0]
    for _ in range(7):
        inv.append(inv[-1] / 1.2)
    sizes = [(int(round(W * s)), int(round(H * s))) for s in inv]
    fast20 = cv2.FastFeatureDetector_create(20, True)
    out = {"pyramid": 0.0, "fast_whole_level": 0.0, "blur": 0.0}
    for _ in range(reps):
        t0 = time.perf_counter()
        levels = [frame]
        for l in range(1, 8):
            levels.append(cv2.resize(levels[-1], sizes[l], interpolation=cv2.INTER_LINEAR))
        padded = [cv2.copyMakeBorder(v, 19, 19, 19, 19, cv2.BORDER_REFLECT_101) for v in levels]
        t1 = time.perf_counter()
        n = sum(len(fast20.detect(v[3:-3, 3:-3])) for v in padded)
        t2 = time.perf_counter()
        for v in levels:
            cv2.GaussianBlur(v, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        t3 = time.perf_counter()
        out["pyramid"] += (t1 - t0) * 1e3 / reps
        out["fast_whole_level"] += (t2 - t1) * 1e3 / reps
        out["blur"] += (t3 - t2) * 1e3 / reps
    out["sum_ms_per_frame"] = out["pyramid"] + out["fast_whole_level"] + out["blur"]
    out["fast_corners_at_20"] = int(n)
    out["threads"] = 1
    out["cv2"] = cv2.__version__
    return out


_JSON_OUT = sys.stdout


def emit_json(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def ctypes_stream(st):
    import ctypes
    return ctypes.c_void_p(st.cuda_stream)


def call_latencies(orbfe, device):
    """Per-call milliseconds of the matcher entry points at SLAM-frame sizes (host arrays in and out,
    as the C++ adapter issues them): C2 stereo pair, C3-sized kNN, local-map projection search."""
    import synth

    def timeit(fn, n=30):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e3
    left, right = synth.stereo_pair(480, 752, 3)
    gl, gr = orbfe.ORBextractor(1200, device=device), orbfe.ORBextractor(1200, device=device)
    _, kl, dl = gl(left, None, (0, 0))
    _, kr, dr = gr(right, None, (0, 0))
    m = orbfe.ORBmatcher(0.8, True, device=device)
    out = {"stereo_match_1200x1200": timeit(lambda: orbfe.ORBmatcher.ComputeStereoMatches(gl, gr, kl, dl, kr, dr, 47.9, 0.11)),
           "knn2_ratio_1200x1200": timeit(lambda: m.knn2(dl, dr))}
    d = synth.map_vs_frame(3000, 1200, 1, w=752, h=480)
    pts = dict(u=d["u"], v=d["v"], ur=d["u"], radius=np.full(3000, 10, np.float32), min_level=np.zeros(3000, np.int32),
               max_level=np.full(3000, -1, np.int32), angle=np.zeros(3000, np.float32), valid=np.ones(3000, np.uint8),
               blocks=np.ones(3000, np.uint8), desc=d["mdesc"])
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], None)
    cl, asg = np.zeros(1200, np.uint8), np.full(1200, -1, np.int32)
    out["search_by_projection_3000pts"] = timeit(lambda: m.SearchByProjection(F, pts, cl, asg))
    out["search_last_frame_3000pts"] = timeit(lambda: m.SearchByProjectionLastFrame(F, pts, cl, asg))
    # keyframe-side searches and the bag-of-words path
    sf = d["scale_factors"]
    kpts = dict(pts, min_level=(d["level"] - 1).astype(np.int32), max_level=d["level"].astype(np.int32))
    inv_s2 = (np.float32(1) / (sf * sf)).astype(np.float32)
    out["fuse_search_3000pts"] = timeit(lambda: m.FuseSearch(F, kpts, inv_s2))
    p12 = dict(u=d["keys"]["x"], v=d["keys"]["y"], radius=np.full(1200, 10, np.float32),
               min_level=(d["keys"]["octave"] - 1).astype(np.int32), max_level=d["keys"]["octave"].astype(np.int32),
               valid=np.ones(1200, np.uint8), desc=d["fdesc"])
    out["search_by_sim3_1200x1200"] = timeit(lambda: m.SearchBySim3(F, F, p12, p12))
    voc = synth.make_vocabulary_fast(10, 6, 1)                    # the shape of ORBvoc.txt: 1.1 M nodes, 10^6 words
    gv = orbfe.ORBVocabulary(10, 6, voc["parent"], voc["desc"], voc["weight"], device=device)
    out["bow_transform_1200feat_k10_L6"] = timeit(lambda: gv.transform_features(dl, 4))
    # batched, device-resident: the descriptors of 1000 frames (1 M) through the ORBvoc-sized tree in one launch
    try:
        import torch
        from orbfe import _lib
        dev = torch.device("cuda", device)
        nbig = 1 << 20
        d_desc = torch.randint(0, 256, (nbig, 32), dtype=torch.uint8, device=dev)
        d_word = torch.empty(nbig, dtype=torch.int32, device=dev)
        d_node = torch.empty(nbig, dtype=torch.int32, device=dev)
        d_w = torch.empty(nbig, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev)

        def big():
            _lib.check(_lib.lib().orbfe_bow_transform_device(gv.h, _lib.ptr(d_desc), nbig, 4, _lib.ptr(d_word), _lib.ptr(d_w),
                                                             _lib.ptr(d_node), ctypes_stream(st)))
        for _ in range(2):
            big()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            big()
        e1.record()
        torch.cuda.synchronize()
        out["bow_transform_1M_desc_device_ms"] = e0.elapsed_time(e1) / 5
    except Exception as e:          # the batched figure is informative only
        out["bow_transform_1M_desc_device_ms"] = "error: " + str(e)[:120]
    _, fvl = gv.transform(dl, 4)
    _, fvr = gv.transform(dr, 4)
    ones_l = np.ones(len(dl), np.uint8)
    out["search_by_bow_1200x1200"] = timeit(lambda: m.SearchByBoW((dl, kl["angle"], ones_l, fvl), (dr, kr["angle"], None, fvr)))
    f12 = np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], np.float32)
    tri = (lambda: m.SearchForTriangulation((kl, dl, None, np.zeros(len(dl), np.uint8), fvl),
                                            (kr, dr, None, np.zeros(len(dr), np.uint8), fvr), f12,
                                            np.array([-1e4, -1e4], np.float32), sf, sf * sf))
    out["search_for_triangulation_1200x1200"] = timeit(tri)
    try:    # fisheye stereo geometry per match (KannalaBrandt8::TriangulateMatches), 1200 ratio-test survivors
        rng = np.random.default_rng(11)
        P1 = np.array([190.978477, 190.973307, 254.931706, 256.897442, 0.003482389, 0.000715034, -0.002053236, 0.000202936], np.float32)
        P2 = np.array([190.442369, 190.434448, 252.598029, 254.917267, 0.003400724, 0.001766232, -0.002663594, 0.000329930], np.float32)
        c1, c2 = orbfe.KannalaBrandt8(P1, device=device), orbfe.KannalaBrandt8(P2, device=device)
        R12, t12 = np.eye(3, dtype=np.float32), np.array([0.1, 0.002, -0.001], np.float32)
        X1 = np.stack([rng.uniform(-2, 2, 1200), rng.uniform(-2, 2, 1200), rng.uniform(0.5, 3, 1200)], 1).astype(np.float32)
        pt1, pt2 = c1.project(X1), c2.project(X1 - t12)
        ones = np.ones(1200, np.float32)
        depth, _ = c1.TriangulateMatches(c2, pt1, pt2, R12, t12, ones, ones)
        assert (depth > 0).mean() > 0.9, "triangulation of exact projections rejected"
        out["kb8_triangulate_1200_matches"] = timeit(lambda: c1.TriangulateMatches(c2, pt1, pt2, R12, t12, ones, ones))
    except Exception as e:          # informative only
        out["kb8_triangulate_1200_matches"] = "error: " + str(e)[:120]
    out["cpu_reference"] = cpu_reference_call_latencies(d, pts, dl, dr, kl, kr, timeit)
    return out


def cpu_reference_call_latencies(d, pts, dl, dr, kl, kr, timeit):
    """The reference's own functions (oracle/_ref/libref_orbmatcher.so: ORBmatcher bodies and DBoW2 compiled from the
    reference tree) on the same inputs, one host thread, as the reference runs them.  Reported beside the GPU call
    latencies; empty when the library was not built."""
    try:
        from oracle import ref as R
        if not R.matcher_available():
            return {}
        import tempfile
        import synth
        out = {}
        R.set_bounds(d["bounds"])
        F = R.RefFrame(d["keys"], d["fdesc"], d["scale_factors"])
        n = len(pts["u"])
        mp = dict(in_view=np.ones(n, np.uint8), proj_x=pts["u"], proj_y=pts["v"], proj_xr=pts["u"], level=d["level"],
                  view_cos=np.full(n, 0.9, np.float32), desc=pts["desc"])
        out["search_by_projection_3000pts"] = timeit(lambda: F.search_mappoints(mp, 1.0, False, 0.0, 0.8), 10)
        voc = synth.make_vocabulary(10, 4, 3)                     # text vocabularies of ORBvoc size are 150 MB: use 10^4 words
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "voc.txt")
            synth.write_vocabulary_text(path, voc)
            rv = R.RefVocabulary(path)
        out["bow_transform_1200feat_k10_L4"] = timeit(lambda: rv.transform(dl, 2), 10)
        sf = d["scale_factors"]
        KF, FR = R.RefFrame(kl, dl, sf), R.RefFrame(kr, dr, sf)
        KF.set_mappoints(np.ones(len(kl), np.uint8))
        R.compute_bow(KF, rv, 2); R.compute_bow(FR, rv, 2)
        out["search_by_bow_1200x1200_k10_L4"] = timeit(lambda: R.search_by_bow_kf_f(KF, FR, 0.8, True), 10)
        return out
    except Exception as e:                                         # the checker is optional for the bench line
        return {"error": str(e)[:200]}


CPU_BUILD = ("reference sources (ORBextractor.cc verbatim) over oracle/cvprims.cpp = scalar restatement of the OpenCV "
             "primitives (bit-exact with cv2 4.13), g++ -O2 -ffp-contract=off; the reference's own build is -O3 -march=native "
             "against OpenCV's SIMD code, see cpu_baseline.cv2_primitives for that part")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = set_workload(args.workload)
    cores = os.cpu_count() or 1
    frames = make_frames(16)
    one, kind = cpu_reference_run(1, 1, frames)                 # calibrate: single-thread frames/s
    per = int(max(2, min(64, round(8.0 * one))))                # ~8 s of CPU work per thread and step
    for _ in range(args.warmup):
        cpu_reference_run(1, cores, frames)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        cpu_reference_run(per, cores, frames)
        tot += per * cores
    dt = time.perf_counter() - t0
    v = tot / dt
    emit_json({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["label"]},
        "run": {"frames_per_step": per * cores},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "build": CPU_BUILD,
                         "sample": f"{per} frames per thread x {cores} threads per step, one ORBextractor per thread",
                         "cv2_primitives": cv2_primitive_ms(frames[0])},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def hbm_peak():
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        return float(json.load(open(peaks_file))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def device_leg(torch, orbfe, dev, local, frames, steps, warmup, barrier, max_over_ranks, world, min_clock_s=0.0):
    """Device-resident throughput of one workload (H, W, NF globals): frames already in HBM, K steps of
    orbfe_extract_batch_device on a stream, CUDA events on that stream.  Returns a dict with the bench figures and the
    per-kernel roofline block."""
    B = len(frames)
    ex = orbfe.ORBextractor(NF, 1.2, 8, 20, 7, device=local)
    cap = ex.capacity
    ex.set_max_bytes(64 << 30)                        # one chunk: the whole batch per launch set
    d_img = torch.from_numpy(frames).to(dev)
    d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.empty(B, dtype=torch.int32, device=dev)
    d_mono = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)

    def step_device():
        ex.extract_batch_device(d_img, LAP, d_kps, d_desc, d_n, d_mono, st)

    for _ in range(max(warmup, 3)):
        step_device()
    st.synchronize()
    geo = ex.frame_geometry()
    n_kp = d_n.cpu().numpy().astype(np.int64)
    cands = sum(len(ex.debug_candidates(l, frame=0)) for l in range(8))
    clocks = ClockSampler(local)
    clocks.start()
    t_wait = time.time()
    while len(clocks.lines) < 6 and time.time() - t_wait < 6.0:  # nvidia-smi needs ~1 s to start; sample under this load
        step_device()
        st.synchronize()
    barrier()
    l0 = ex.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        step_device()
    e1.record(st)
    st.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ex.launch_count() - l0
    # per-kernel times: a second, profiled pass (CUDA events between the kernels on the launching stream), so that the
    # timed pass above carries no event records.  (Running the blur on a side stream beside the latency-bound quadtree
    # was measured: 6.03 vs 6.07 ms per step, i.e. the two kernels do not overlap usefully -- not kept.)
    ex.set_profiling(True)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(st)
    nprof = max(3, min(steps, 10))
    for _ in range(nprof):
        step_device()
    e3.record(st)
    st.synchronize()
    stage = ex.stage_ms()
    ms_serial = e2.elapsed_time(e3) / nprof
    ex.set_profiling(False)
    # the clock sampler keeps running over the same load until it has seen `min_clock_s` of it (untimed steps)
    t_more = time.time()
    while time.time() - t_more < min_clock_s:
        step_device()
        st.synchronize()
    clk = clocks.stop()
    value = world * B * steps / (ms / 1e3)

    sizes = [ex.level_size(l, (H, W)) for l in range(8)]
    P = sum(w * h for w, h in sizes)
    P06 = sum(w * h for w, h in sizes[:7])
    Ppad = sum((w + 38) * (h + 38) for w, h in sizes)
    K = float(n_kp.mean())
    C = float(cands)
    alg = {"pyramid": P06 + Ppad, "fast": P + 12 * C, "octree": 12 * C + 12 * K,
           "layout": 28 * K, "blur": 2 * P, "describe": (749 + 4 + 512 + 32) * K}
    frame_bytes = 3 * P + P06 + Ppad + 24 * C + 1337 * K
    kern = {k: v for k, v in stage.items() if k in alg}
    top = max(kern, key=kern.get)
    peak, peak_src = hbm_peak()
    achieved = alg[top] * B / (kern[top] / 1e3) / 1e9
    traffic, traffic_src = None, None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tfile):
        ent = json.load(open(tfile)).get(f"{H}x{W}", {}).get(top, {})
        if ent.get("dram_bytes_per_frame") is not None:
            traffic = ent["dram_bytes_per_frame"] * B
            traffic_src = ent.get("source")
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[top] * B, "launch_ms": kern[top],
                "kernel_ms_per_step": kern, "kernel_ms_sum": sum(kern.values()), "ms_per_step_kernels_back_to_back": ms_serial,
                "kernel_ms_how": "separate profiled pass of the same steps (CUDA events between the kernels)",
                "pipeline_bytes_per_frame": frame_bytes,
                "pipeline_frac": value / world * frame_bytes / 1e9 / peak}
    return dict(ex=ex, value=value, ms=ms, launches=launches, clk=clk, geo=geo, n_kp=n_kp, K=K, C=C, cap=cap,
                roofline=roofline)


def pairs_leg(torch, orbfe, dev, local, kind, B, steps, barrier, max_over_ranks, world, rank):
    """BASELINE configs 2 / 3 as batches of independent stereo pairs (frame pairs shard over the ranks, no collective):
    c2 = 752x480 rectified pairs, nFeatures 1200, lapping {0,0}: left + right extraction + Frame::ComputeStereoMatches;
    c3 = 512x512 fisheye-style pairs, nFeatures 1500, lapping {0,511}: left + right extraction + the kNN-2 / 0.7 ratio
    matcher of Frame::ComputeStereoFishEyeMatches.  Device-resident (images in HBM) and end to end (pinned host images
    in, keypoints / descriptors / matches out, copies inside the timed region)."""
    import synth
    from orbfe import _lib
    if kind == "c2":
        h, w, nf, lap, gen = 480, 752, 1200, (0, 0), synth.stereo_pair
        label = "C2 752x480 x2 rectified pairs, nFeatures=1200, L/R extraction + ComputeStereoMatches"
    else:
        h, w, nf, lap, gen = 512, 512, 1500, (0, 511), synth.shifted_pair
        label = "C3 512x512 x2 fisheye-style pairs, nFeatures=1500, lapping {0,511}, L/R extraction + kNN-2 + 0.7 ratio"
    base = [gen(h, w, 11 + i + 100 * rank) for i in range(8)]
    rng = np.random.default_rng(3 + rank)
    L, R = np.empty((B, h, w), np.uint8), np.empty((B, h, w), np.uint8)
    for i in range(B):
        a, b = base[i % len(base)]
        k = 0 if i < len(base) else int(rng.integers(1, h))       # the same row shift on both images keeps the geometry
        L[i], R[i] = np.roll(a, k, axis=0), np.roll(b, k, axis=0)
    hL, hR = torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()
    exL, exR = orbfe.ORBextractor(nf, device=local), orbfe.ORBextractor(nf, device=local)
    for ex in (exL, exR):
        ex.set_max_bytes(64 << 30)
    cap = exL.capacity
    st = torch.cuda.Stream(device=dev)
    mk = lambda *shape, dt=torch.uint8: torch.empty(shape, dtype=dt, device=dev)
    dL, dR = hL.to(dev), hR.to(dev)
    oL = (mk(B, cap, 28), mk(B, cap, 32), mk(B, dt=torch.int32), mk(B, dt=torch.int32))
    oR = (mk(B, cap, 28), mk(B, cap, 32), mk(B, dt=torch.int32), mk(B, dt=torch.int32))
    mbf, mb = 47.9, 47.9 / 435.2
    res = {}

    def step(imgL, imgR):
        exL.extract_batch_device(imgL, lap, oL[0], oL[1], oL[2], oL[3], st)
        exR.extract_batch_device(imgR, lap, oR[0], oR[1], oR[2], oR[3], st)
        if kind == "c2":
            res["m"] = orbfe.ORBmatcher.ComputeStereoMatchesBatchDevice(exL, exR, oL[0], oL[1], oL[2], oR[0], oR[1], oR[2], mbf, mb, st)
        else:
            res["m"] = orbfe.ORBmatcher.knn2_batch_device(oL[1], oL[3], oL[2], oR[1], oR[3], oR[2], st)
    for _ in range(3):
        step(dL, dR)
    st.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        step(dL, dR)
    e1.record(st)
    st.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    if kind == "c2":
        matched = float((res["m"][0] >= 0).sum().item()) / B
    else:
        matched = float((res["m"][2] >= 0).sum().item()) / B
    # end to end: pinned host images in, results out, every step.  Three streams (H2D, compute, D2H) and two sets of
    # device buffers: the images of step k + 1 arrive and the results of step k - 1 leave while step k computes -- every
    # step still copies its own inputs in and its own results out.  `sync_ms` beside it is one blocking step at a time.
    sH, sD = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ins = [(torch.empty_like(dL), torch.empty_like(dR)) for _ in range(2)]
    outs = [(tuple(torch.empty_like(t) for t in oL), tuple(torch.empty_like(t) for t in oR)) for _ in range(2)]
    host_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (oL[0], oL[1], oL[2], oR[0], oR[1], oR[2])] for _ in range(2)]
    host_m = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in res["m"]] for _ in range(2)]

    def step_set(b):
        (iL, iR), (pL, pR) = ins[b], outs[b]
        with torch.cuda.stream(st):
            exL.extract_batch_device(iL, lap, pL[0], pL[1], pL[2], pL[3], st)
            exR.extract_batch_device(iR, lap, pR[0], pR[1], pR[2], pR[3], st)
            if kind == "c2":
                return orbfe.ORBmatcher.ComputeStereoMatchesBatchDevice(exL, exR, pL[0], pL[1], pL[2], pR[0], pR[1], pR[2], mbf, mb, st)
            return orbfe.ORBmatcher.knn2_batch_device(pL[1], pL[3], pL[2], pR[1], pR[3], pR[2], st)

    def pipeline(k_steps, overlap):
        evC, evD, keep = [None, None], [None, None], [None, None]
        for k in range(k_steps):
            b = k & 1
            if evC[b] is not None:
                sH.wait_event(evC[b])              # step k - 2 has read this input set
            with torch.cuda.stream(sH):
                ins[b][0].copy_(hL, non_blocking=True)
                ins[b][1].copy_(hR, non_blocking=True)
                evH = torch.cuda.Event()
                evH.record(sH)
            st.wait_event(evH)
            if evD[b] is not None:
                st.wait_event(evD[b])              # the results of step k - 2 have left this output set
            keep[b] = step_set(b)
            evC[b] = torch.cuda.Event()
            evC[b].record(st)
            sD.wait_event(evC[b])
            with torch.cuda.stream(sD):
                pL, pR = outs[b]
                for d, s_ in zip(host_out[b], (pL[0], pL[1], pL[2], pR[0], pR[1], pR[2])):
                    d.copy_(s_, non_blocking=True)
                for d, s_ in zip(host_m[b], keep[b]):
                    d.copy_(s_, non_blocking=True)
                evD[b] = torch.cuda.Event()
                evD[b].record(sD)
            if not overlap:
                sD.synchronize()
        sD.synchronize()
        st.synchronize()

    def timed(overlap):
        pipeline(2, overlap)
        barrier()
        t0 = time.perf_counter()
        pipeline(steps, overlap)
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt * 1e3) / steps
    sync_ms = timed(False)
    e2e_ms = timed(True)
    # the pipelined results equal the resident ones (same inputs)
    torch.cuda.synchronize()
    k_last = (steps - 1) & 1
    assert torch.equal(host_out[k_last][2], oL[2].cpu()) and torch.equal(host_m[k_last][0], res["m"][0].cpu()), "pipelined pairs differ"
    host_out, host_m = host_out[0], host_m[0]
    h2d = 2 * B * h * w
    d2h = sum(t.numel() * t.element_size() for t in host_out + host_m)
    return {"metric": "stereo pairs/s (" + label + ")", "value": world * B / (ms / 1e3), "unit": "pairs/s", "n_gpus": world,
            "ms_per_step": ms, "pairs_per_gpu_per_step": B, "matches_per_pair": matched,
            "keypoints_per_frame": float(oL[2].float().mean().item()),
            "e2e": {"value": world * B / (e2e_ms / 1e3), "unit": "pairs/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "sync_step_value": world * B / (sync_ms / 1e3), "sync_step_ms": sync_ms,
                    "how": "H2D of both image batches, 2 extractions + matcher, D2H of all results per step; three streams and two buffer sets, so the copies of neighbouring steps overlap the kernels (sync_step_value: one blocking step at a time)"},
            "residency": "value: images resident in HBM, CUDA events on the launching stream, max over ranks"}


def copy_ceiling(torch, dev, h2d_bytes, d2h_bytes, reps, barrier, max_over_ranks):
    """What the host link alone allows: the step's H2D and D2H volumes copied concurrently on two streams from / into
    pinned memory, no kernels; with N ranks all ranks copy at the same time (one host, shared root complex / memory)."""
    src = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(src, non_blocking=True)
        with torch.cuda.stream(s2):
            dst.copy_(d_out, non_blocking=True)
    once()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return max_over_ranks(dt * 1e3) / reps


def bind_to_gpu_numa(local):
    """Pin this rank's threads (and with them its first-touch pinned allocations) to the cores of its GPU's NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"cpus": len(cpus), "first": cpus[0] if cpus else None, "last": cpus[-1] if cpus else None}
    except Exception as e:
        return {"error": str(e)[:100]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import orbfe
    from orbfe import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    orbfe.lib()
    wl = set_workload(args.workload)
    B = args.frames or wl["frames"]
    frames = make_frames(B, seed0=7 + 1000 * rank)
    pinned = torch.from_numpy(frames).pin_memory()

    # ---- device-resident throughput (value) ----
    leg = device_leg(torch, orbfe, dev, local, frames, args.steps, args.warmup, barrier, max_over_ranks, world, min_clock_s=2.0)
    ex, value, ms, launches, clk, geo, n_kp, K, C, cap = (leg[k] for k in ("ex", "value", "ms", "launches", "clk", "geo", "n_kp", "K", "C", "cap"))
    roofline = leg["roofline"]

    # ---- end to end through the host-pointer C ABI (e2e) ----
    ex.set_max_bytes(int(geo["per_frame_bytes"]) * args.chunk)   # pipeline H2D / kernels / D2H per chunk
    out = (torch.empty(B, dtype=torch.int32).pin_memory().numpy(), torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
           torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory().numpy().view(_lib.KP_DTYPE).reshape(B, cap),
           torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory().numpy())
    for _ in range(2):
        ex.extract_batch(pinned, LAP, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ex.extract_batch(pinned, LAP, out=out)        # synchronous: returns with results in host memory
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    e2e_sync_ms = max_over_ranks(dt * 1e3)
    e2e_sync = world * B * args.steps / (e2e_sync_ms / 1e3)
    assert np.array_equal(out[0].astype(np.int64), n_kp), "host and device paths disagree"
    # streaming form of the same call (orbfe_extract_batch_submit / _wait), two batches in flight: step k+1 is
    # submitted before step k is waited for, so its H2D runs under the kernels of step k.  Every step still copies
    # its own frames in and its own results out, from / into alternating pinned buffers.
    ins = (pinned, pinned.clone().pin_memory())
    outs = (out, tuple(torch.from_numpy(a.view(np.uint8)).clone().pin_memory().numpy().view(a.dtype).reshape(a.shape) for a in out))
    for a in outs[1]:
        a.view(np.uint8)[...] = 0

    def stream_steps(k):
        ex.extract_batch_submit(ins[0], LAP, outs[0])
        for i in range(1, k):
            ex.extract_batch_submit(ins[i & 1], LAP, outs[i & 1])
            ex.extract_batch_wait()
        ex.extract_batch_wait()
    stream_steps(2)
    barrier()
    t0 = time.perf_counter()
    stream_steps(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    e2e_ms = max_over_ranks(dt * 1e3)
    e2e = world * B * args.steps / (e2e_ms / 1e3)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and \
        np.array_equal(outs[0][0].astype(np.int64), n_kp), "streaming results differ"
    for f in range(B):
        k = int(outs[0][0][f])
        assert outs[0][2][f, :k].tobytes() == outs[1][2][f, :k].tobytes() and \
            np.array_equal(outs[0][3][f, :k], outs[1][3][f, :k]), "streaming results differ"
    assert np.array_equal(out[0].astype(np.int64), n_kp), "host and device paths disagree"
    h2d = B * H * W
    d2h = B * (cap * 28 + cap * 32 + 8)
    # the host link's own ceiling for these copy volumes, all ranks copying at once (no kernels)
    ceil_ms = copy_ceiling(torch, dev, h2d, d2h, max(3, min(args.steps, 10)), barrier, max_over_ranks)
    ceil_fps = world * B / (ceil_ms / 1e3)

    # ---- one frame per call, as Frame::ExtractORB issues it (latency, not throughput) ----
    one = frames[0]
    for _ in range(5):
        ex(one, None, LAP)
    t0 = time.perf_counter()
    for _ in range(50):
        ex(one, None, LAP)
    single_ms = (time.perf_counter() - t0) / 50 * 1e3
    latency = {"extract_1_frame": single_ms}
    if not args.no_match and rank == 0:
        latency.update(call_latencies(orbfe, local))

    # ---- matching leg: 2000 frame descriptors vs 1 M map descriptors, map sharded over ranks ----
    matching = None
    if not args.no_match:
        matching = matching_leg(torch, dist, orbfe, dev, rank, world, args.steps, barrier, max_over_ranks)

    # ---- the other throughput workload of BASELINE.json (C4 when the line is C1), device-resident, same method ----
    extra = None
    other = "c4" if args.workload == "c1" else None
    if other and not args.no_extra:
        del ex, leg
        torch.cuda.empty_cache()
        owl = set_workload(other)
        oframes = make_frames(args.extra_frames or owl["frames"], seed0=7 + 1000 * rank)
        oleg = device_leg(torch, orbfe, dev, local, oframes, max(3, min(args.steps, 10)), 3, barrier, max_over_ranks, world)
        extra = {other: {"metric": METRIC, "value": oleg["value"], "unit": "frames/s", "n_gpus": world,
                         "ms_per_step": oleg["ms"] / max(3, min(args.steps, 10)), "workload": owl["label"],
                         "frames_per_gpu_per_step": len(oframes), "keypoints_per_frame": oleg["K"],
                         "fast_candidates_per_frame": oleg["C"], "clocks": oleg["clk"], "roofline": oleg["roofline"],
                         "residency": "device-resident (frames in HBM), CUDA events on the launching stream, max over ranks"}}
        set_workload(args.workload)
        del oleg
        torch.cuda.empty_cache()
        for kind in ("c2", "c3"):
            try:
                extra[kind] = pairs_leg(torch, orbfe, dev, local, kind, 256, max(3, min(args.steps, 10)), barrier, max_over_ranks,
                                        world, rank)
            except Exception as e:      # reported, never fatal for the headline
                extra[kind] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the reference's ORBextractor on the host cores (bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        one_t, kind = cpu_reference_run(1, 1, frames)
        per = int(max(2, min(64, round(12.0 * one_t))))
        v, kind = cpu_reference_run(per, cores, frames)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "single_thread_frames_per_s": one_t,
               "build": CPU_BUILD,
               "sample": f"{per} frames per thread x {cores} threads of the same workload, one ORBextractor per thread",
               "cv2_primitives": cv2_primitive_ms(frames[0])}

    emit_json({
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["label"]},
        "run": {"frames_per_gpu_per_step": B, "keypoints_per_frame": K, "fast_candidates_per_frame": C,
                "l2": f"inputs+intermediates per step = {B * geo['per_frame_bytes'] / 1e6:.0f} MB > 126 MB L2 (no flush needed)",
                "e2e_chunk_frames": args.chunk, "e2e_timer": "host perf_counter around the K steps, max over ranks",
                "e2e_mode": "orbfe_extract_batch_submit/_wait, 2 host batches in flight (sync_call_value: one blocking orbfe_extract_batch per step)",
                "cpu_affinity": affinity,
                "ncu_captures": "profiles/*_launches_1024.txt = launch list of this command at 1024 frames per launch (kernel shares of the step); --set full captures run at --frames 128 (per-frame counters, not absolute times)"},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "sync_call_value": e2e_sync,
                "sync_call_ms_per_step": e2e_sync_ms / args.steps, "single_frame_call_ms": single_ms,
                "copy_ceiling_frames_per_s": ceil_fps, "copy_ceiling_ms_per_step": ceil_ms,
                "frac_of_ceiling": e2e / ceil_fps,
                "copy_ceiling_how": "the step's H2D and D2H bytes copied concurrently on two streams, pinned host memory, "
                                    "no kernels, all ranks at once; device-resident `value` is the other bound"},
        "call_latency_ms": latency,
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "matching": matching,
        "extra": extra,
    })
    if world > 1:
        dist.destroy_process_group()


def matching_leg(torch, dist, orbfe, dev, rank, world, steps, barrier, max_over_ranks):
    """C5: 2000 frame descriptors against a 1 M descriptor map sharded over the ranks.  Brute-force kNN-2 (exchange of
    the per-shard best two fused into the merge kernel over NVLink peer loads) and the map-sharded SearchByProjection."""
    import orbfe.dist as D
    nq, nmap = 2000, 1000000
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    lo, hi = rank * nmap // world, (rank + 1) * nmap // world
    shard = np.random.default_rng(100 + rank).integers(0, 256, (hi - lo, 32), dtype=np.uint8)
    d_q = torch.from_numpy(q).to(dev)

    def time_exchange(exchange):
        smap = D.ShardedMap(shard, lo, dev, exchange=exchange)     # this rank's shard, resident in HBM
        for _ in range(4):
            out = smap.knn2(d_q)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(steps):
            out = smap.knn2(d_q)                   # local kNN-2 -> exchange of the per-shard best two -> merge
        m1.record()
        torch.cuda.synchronize()
        barrier()
        eager = max_over_ranks(m0.elapsed_time(m1))
        res = [t.clone() for t in out]
        # The step is a few hundred microseconds of device work issued from Python (tensor allocations, ctypes): the
        # eager loop above measures the interpreter as much as the GPU.  The same two steps (both symmetric buffers)
        # captured once into a CUDA graph and replayed give the device-side rate; both are reported.
        graph_ms = None
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                smap.knn2(d_q)
                gout = smap.knn2(d_q)
            reps = max(1, steps // 2)
            g.replay()
            torch.cuda.synchronize()
            barrier()
            m0.record()
            for _ in range(reps):
                g.replay()
            m1.record()
            torch.cuda.synchronize()
            barrier()
            ok = all(torch.equal(a_, b_) for a_, b_ in zip(res, gout))
            ok = max_over_ranks(0.0 if ok else 1.0) == 0.0
            if ok:
                graph_ms = max_over_ranks(m0.elapsed_time(m1)) * steps / (2 * reps)
        except Exception as e:      # noqa: BLE001 -- capture not possible on this setup: the eager number stands
            sys.stderr.write("bench: CUDA graph capture of the kNN step failed (%s); eager timing only\n" % str(e)[:120])
            torch.cuda.synchronize()
        return (graph_ms if graph_ms is not None else eager), smap.exchange, res, eager, graph_ms is not None
    # exchange fused into the merge kernel (peer loads over NVLink, symmetric memory); all-gather form beside it.
    # Both forms are timed twice in alternation and the faster pass of each is reported (sub-millisecond steps).
    mms, how, res, eager_ms, graphed = time_exchange("p2p")
    nms = None
    if world > 1:
        nms, _, res2, _, _ = time_exchange("nccl")
        assert all(torch.equal(a, b) for a, b in zip(res, res2)), "p2p and all-gather exchanges disagree"
        r2 = time_exchange("p2p")
        if r2[0] < mms:
            mms, eager_ms, graphed = r2[0], r2[3], r2[4]
        nms = min(nms, time_exchange("nccl")[0])
    pairs = nq * nmap * steps / (mms / 1e3)
    # Roofline of the brute force.  The kernel is csrc/knn_umma.cu: one signed byte per descriptor
    # bit, S = 256 - 2 * hamming from tcgen05.mma kind::i8, 2 * 256 integer operations per descriptor pair.  Peak: the
    # dense int8 rate of the tensor cores = 2 x the dense bf16 rate the driver measured on this pool (MEASURED_PEAKS.json
    # has no int8 figure; nominal 4500 TOP/s).  Under it sits the read of the accumulators: 4 B of tensor memory per pair
    # at 64 B / clk / SM (B300_MICROARCH.md, "LDTM throughput"), which the packed 16-bit tcgen05.ld halves on the
    # register side.  The scalar kernel it replaces ran at 0.68 T pairs/s = 0.46 of the POPC pipe (profiles/r2_match_ncu_summary.txt).
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16 = float(json.load(open(peaks_file)).get("bf16_tflops", 0.0)) if os.path.exists(peaks_file) else 0.0
    int8_peak, peak_src = (2.0 * bf16, "2 x MEASURED_PEAKS.json bf16_tflops (dense int8 = 2 x dense bf16)") if bf16 > 0 else (4500.0, "nominal dense int8, 4500 TOP/s")
    tops = pairs * 512.0 / 1e12
    matching = {"metric": "Hamming matches/s (2000 frame x 1M map descriptors, kNN-2 + ratio)", "value": pairs,
                "unit": "descriptor pairs/s", "ms_per_step": mms / steps, "map_shards": world, "gather": "none",
                "timing": ("two steps captured in a CUDA graph and replayed (device-side rate)" if graphed
                           else "eager Python calls"),
                "ms_per_step_eager_python_calls": eager_ms / steps,
                "roofline": {"bound": "tensor", "unit": "TOP/s (int8)", "kernel": "k_knn2_umma", "peak": int8_peak * world,
                             "peak_source": peak_src, "achieved": tops, "frac": tops / (int8_peak * world),
                             "ops_per_pair": 512,
                             "accumulator_read": {"bytes_per_pair": 4, "achieved_GBs": pairs * 4.0 / 1e9,
                                                  "peak_GBs": 64.0 * 148 * 1965e6 * world / 1e9,
                                                  "frac": pairs * 4.0 / (64.0 * 148 * 1965e6 * world),
                                                  "note": "peak = 64 B/clk/SM of 32-bit tcgen05.ld; the kernel reads with .pack::16b (two columns per register), which is how it can sit above 1.0"},
                             # north_star's INT view of the same number: the plain statement of DescriptorDistance needs
                             # 8 POPC32 per pair; the POPC pipe (25 lanes/clk/SM measured, profiles/r1_pipe_bench.txt) caps a
                             # popcount kernel at peak / 8 pairs/s (the scalar kernel reached 0.46 of it with 5 POPC per pair)
                             "int_view": {"algorithmic_popc32_per_s": pairs * 8.0, "popc_pipe_peak_per_s": 25.0 * 148 * 1965e6 * world,
                                          "ratio_to_popc_pipe_peak": pairs * 8.0 / (25.0 * 148 * 1965e6 * world)},
                             "scalar_kernel_pairs_per_s_1gpu": 0.68e12,
                             "ncu": "profiles/r2_match_ncu_summary.txt"}}
    if world > 1:
        matching["gather"] = ("peer loads inside the merge kernel (symmetric memory over NVLink) + 1 device barrier"
                              if how == "p2p" else how)
        matching["ms_per_step_nccl_all_gather"] = nms / steps
    try:
        matching["search_by_projection"] = D.bench_sharded_projection(torch, dist, orbfe, dev, rank, world, steps,
                                                                      barrier, max_over_ranks)
    except Exception as e:     # reported, never fatal for the extraction line
        matching["search_by_projection"] = {"error": str(e)[:200]}
    return matching


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS), help="BASELINE.json config the line is measured on")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: 1024 for c1, 384 for c4)")
    ap.add_argument("--extra-frames", type=int, default=0, help="frames per GPU per step of the extra C4 block")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra block (C4 device-resident throughput)")
    ap.add_argument("--chunk", type=int, default=256, help="frames per pipelined chunk on the host-pointer path")
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: keep a private handle to it and point fd 1 at stderr, so that anything a
    # library prints on stdout (NCCL's version banner under NCCL_DEBUG=VERSION, for one) cannot end up beside it
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    _JSON_OUT.flush()


if __name__ == "__main__":
    main()
