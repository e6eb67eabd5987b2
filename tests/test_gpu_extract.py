"""GPU parity: the CUDA extraction path (through the C ABI) against the CPU oracle, stage by
stage and end to end, on the BASELINE configs' shapes.  Bit-exact everywhere (pyramid bytes, FAST
candidate lists, retained keypoints, angles, blurred bytes, descriptors, output order)."""
import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

CONFIGS = [  # (h, w, nfeatures, lapping)  -- BASELINE.json configs C1, C2, C3 (two lapping ranges), C4
    (480, 752, 1000, (0, 1000)),
    (480, 752, 1200, (0, 0)),
    (512, 512, 1500, (0, 511)),
    (512, 512, 1500, (100, 411)),
    (720, 1280, 2000, (0, 1000)),
]


@pytest.fixture(scope="module")
def orbfe():
    import orbfe as m
    m.lib()
    return m


def _check_frame(ex_gpu, ex_cpu, img, lap, stages=True):
    mono_o, kps_o, desc_o = ex_cpu(img, lap)
    mono_g, kps_g, desc_g = ex_gpu(img, None, lap)
    if stages:
        for lvl in range(ex_cpu.nlevels):
            L = ex_cpu.level(lvl)
            assert np.array_equal(ex_gpu.pyramid_level(lvl, with_border=True), L["padded"]), f"pyramid level {lvl}"
            assert np.array_equal(ex_gpu.pyramid_level(lvl), L["padded"][19:-19, 19:-19])
            cg = ex_gpu.debug_candidates(lvl)
            assert np.array_equal(cg, L["cands"]), f"FAST candidates level {lvl}: {len(cg)} vs {len(L['cands'])}"
            kg = ex_gpu.debug_level_keypoints(lvl)
            ko = np.stack([L["kps"]["x"] - 16, L["kps"]["y"] - 16, L["kps"]["response"]], 1).astype(np.int32)
            assert np.array_equal(kg, ko), f"octree level {lvl}"
            if len(L["kps"]):
                assert np.array_equal(ex_gpu.debug_blurred(lvl), L["blurred"]), f"blur level {lvl}"
    assert mono_g == mono_o
    assert len(kps_g) == len(kps_o)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps_g[f], kps_o[f]), f
    assert np.array_equal(kps_g["angle"].view(np.uint32), kps_o["angle"].view(np.uint32)), "angles bit-exact"
    bad = np.flatnonzero((desc_g != desc_o).any(axis=1))
    assert len(bad) <= 1e-3 * len(desc_o), f"{len(bad)} of {len(desc_o)} descriptors differ"
    return len(bad)


@pytest.mark.parametrize("h,w,nf,lap", CONFIGS)
def test_extract_parity_stagewise(orbfe, h, w, nf, lap):
    ex_g, ex_c = orbfe.ORBextractor(nf), O.Extractor(nf)
    assert np.array_equal(ex_g.GetScaleFactors(), ex_c.tables()["scale"])
    assert np.array_equal(ex_g.GetInverseScaleFactors(), ex_c.tables()["inv_scale"])
    assert np.array_equal(ex_g.GetScaleSigmaSquares(), ex_c.tables()["sigma2"])
    assert np.array_equal(ex_g.GetInverseScaleSigmaSquares(), ex_c.tables()["inv_sigma2"])
    assert np.array_equal(ex_g.features_per_level(), ex_c.tables()["nfeatures"])
    nbad = 0
    for seed in (0, 1):
        nbad += _check_frame(ex_g, ex_c, synth.synth_frame(h, w, seed), lap)
    assert nbad == 0, "descriptor mismatches on these seeds would need an explanation"


def test_noise_frames_and_odd_sizes(orbfe):
    for (h, w, nf, seed) in [(200, 260, 300, 3), (241, 323, 50, 4), (480, 640, 3000, 5), (333, 777, 777, 6)]:
        img = synth.noise_frame(h, w, seed) if seed % 2 else synth.synth_frame(h, w, seed)
        _check_frame(orbfe.ORBextractor(nf), O.Extractor(nf), img, (0, 0))


def test_very_wide_frame_more_roots_than_features(orbfe):
    # 9 quadtree roots per level (ORBextractor.cc:718) with a per-level target of ~20 features: the levels return up to
    # 4 nodes per root, more than orbfe_max_keypoints() allows for -- the wrappers size the outputs with
    # orbfe_max_keypoints_for (found by tools/parity_stress.py)
    from orbfe._lib import ptr
    import ctypes as C
    ex_g, ex_c = orbfe.ORBextractor(50, 1.2, 3), O.Extractor(50, 1.2, 3)
    img = synth.noise_frame(175, 1373, 468400454)
    assert ex_g.capacity_for(175, 1373) > ex_g.capacity
    _check_frame(ex_g, ex_c, img, (0, 1000))
    n = C.c_int(0)
    kps = np.empty(ex_g.capacity, orbfe.KP_DTYPE)
    desc = np.empty((ex_g.capacity, 32), np.uint8)
    rc = orbfe.lib().orbfe_extract(ex_g.handle, ptr(img), 175, 1373, 1373, 0, 1000, ptr(kps), ptr(desc),
                                   ex_g.capacity, C.byref(n))
    assert rc == orbfe.ERR_CAPACITY       # reported, never truncated silently


def test_descriptor_on_an_angle_where_libm_sinf_is_not_correctly_rounded(orbfe):
    # tools/parity_stress.py, seed 5 case 60: keypoint 165 of this frame has sinf(angle) 0.517 ulp off in glibc; with the
    # rounded fp64 value a BRIEF sample moved from row 3 to row 4 and bit 198 of its descriptor flipped.  k_describe
    # restates glibc's routine (csrc/sincosf_core.h), so every descriptor is equal.
    img = synth.synth_frame(630, 1185, 75409971)
    assert _check_frame(orbfe.ORBextractor(4000, 1.2, 7, 12, 7), O.Extractor(4000, 1.2, 7, 12, 7), img, (0, 1000), stages=False) == 0


def test_other_pyramid_parameters(orbfe):
    img = synth.synth_frame(480, 640, 9)
    for (nf, sf, nl, ini, mn) in [(800, 2.0, 3, 20, 7), (600, 1.5, 4, 30, 10), (500, 1.1, 10, 12, 5), (400, 1.2, 1, 20, 7)]:
        ex_g = orbfe.ORBextractor(nf, sf, nl, ini, mn)
        ex_c = O.Extractor(nf, sf, nl, ini, mn)
        _check_frame(ex_g, ex_c, img, (0, 0))


def test_strided_input_and_reuse_across_sizes(orbfe):
    ex_g, ex_c = orbfe.ORBextractor(700), O.Extractor(700)
    big = synth.synth_frame(500, 800, 2)
    view = big[10:490, 20:772]          # non-contiguous rows: step != cols
    _check_frame(ex_g, ex_c, view, (0, 1000), stages=False)
    _check_frame(ex_g, ex_c, synth.synth_frame(300, 400, 3), (0, 0), stages=False)   # same handle, new geometry
    _check_frame(ex_g, ex_c, view, (0, 1000), stages=False)


def test_empty_and_error_paths(orbfe):
    ex = orbfe.ORBextractor(500)
    mono, k, d = ex(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(k) == 0
    import ctypes as C
    n = C.c_int(0)
    rc = orbfe.lib().orbfe_extract(ex.handle, None, 0, 0, 0, 0, 0, None, None, 0, C.byref(n))
    assert rc == orbfe.EMPTY_IMAGE                     # ORBextractor.cc:1561-1562
    for shape in [(2, 2), (90, 120)]:                  # a level smaller than the 16-px border window:
        with pytest.raises(orbfe.OrbfeError):          # the reference itself aborts there (negative nIni)
            ex(np.zeros(shape, np.uint8))
    tiny = synth.noise_frame(150, 200, 1)              # top levels have no 35-px FAST cell: no keypoints there
    mo, ko, do = O.Extractor(500)(tiny, (0, 0))
    mg, kg, dg = ex(tiny, None, (0, 0))
    assert mg == mo and kg.tobytes() == ko.tobytes() and np.array_equal(dg, do) and len(kg) > 20
    flat = np.full((480, 752), 97, np.uint8)           # no corners anywhere
    mono, k, d = ex(flat)
    assert mono == 0 and len(k) == 0 and d.shape == (0, 32)


def test_batch_host_and_device_paths(orbfe):
    import torch
    nf, h, w, B = 1000, 480, 752, 6
    frames = np.stack([synth.synth_frame(h, w, 20 + i) for i in range(B)])
    ex_g, ex_c = orbfe.ORBextractor(nf), O.Extractor(nf)
    ref = [ex_c(frames[i], (0, 1000)) for i in range(B)]
    ex_g.set_max_bytes(64 << 20)                       # forces several chunks: exercises the pipeline
    pinned = torch.from_numpy(frames).pin_memory()
    n, mono, kps, desc = ex_g.extract_batch(pinned, (0, 1000))
    for i in range(B):
        assert mono[i] == ref[i][0] and n[i] == len(ref[i][1])
        assert kps[i, :n[i]].tobytes() == ref[i][1].tobytes()
        assert np.array_equal(desc[i, :n[i]], ref[i][2])
    # device-resident path on a torch stream
    ex_g.set_max_bytes(6 << 30)
    dev = torch.device("cuda:0")
    d_img = pinned.to(dev)
    cap = ex_g.capacity
    d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.empty(B, dtype=torch.int32, device=dev)
    d_mono = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ex_g.extract_batch_device(d_img, (0, 1000), d_kps, d_desc, d_n, d_mono, st)
    st.synchronize()
    n2 = d_n.cpu().numpy()
    k2 = d_kps.cpu().numpy().view(orbfe.KP_DTYPE).reshape(B, cap)
    de2 = d_desc.cpu().numpy()
    for i in range(B):
        assert n2[i] == len(ref[i][1]) and int(d_mono[i]) == ref[i][0]
        assert k2[i, :n2[i]].tobytes() == ref[i][1].tobytes()
        assert np.array_equal(de2[i, :n2[i]], ref[i][2])
    assert ex_g.launch_count() > 0


def test_octree_kernel_alone(orbfe):
    ex = orbfe.ORBextractor(1000)
    for seed in range(12):
        rng = np.random.default_rng(seed)
        W, H = int(rng.integers(100, 1300)), int(rng.integers(100, 700))
        if W < H // 2 + 1:
            W = H
        n = int(rng.integers(1, min(9000, W * H // 4)))
        pos = rng.choice(W * H, size=n, replace=False)
        xys = np.stack([pos % W, pos // W, rng.integers(7, 40, n)], 1).astype(np.int32)
        xys = xys[np.lexsort((xys[:, 0], xys[:, 1]))]
        N = int(rng.integers(1, 700))
        assert np.array_equal(ex.debug_octree(xys, 16, 16 + W, 16, 16 + H, N), O.octree(xys, 16, 16 + W, 16, 16 + H, N))


def test_two_instances_from_two_threads(orbfe):
    """The reference runs the left and right extractors concurrently (Frame.cc:136-141)."""
    import threading
    left, right = synth.stereo_pair(480, 752, 4)
    exs = [orbfe.ORBextractor(1200), orbfe.ORBextractor(1200)]
    outs = [None, None]

    def run(i, img):
        for _ in range(3):
            outs[i] = exs[i](img, None, (0, 0))
    ts = [threading.Thread(target=run, args=(0, left)), threading.Thread(target=run, args=(1, right))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i, img in enumerate((left, right)):
        mo, ko, do = O.Extractor(1200)(img, (0, 0))
        assert outs[i][0] == mo and outs[i][1].tobytes() == ko.tobytes() and np.array_equal(outs[i][2], do)


def test_c4_batch_1280x720(orbfe):
    """C4 shape: a batch of 1280x720 frames, nFeatures 2000, through the batched host API in two
    chunks; every frame equals the oracle (4096 frames at full size run in bench.py --frames)."""
    B = 6
    frames = np.stack([synth.synth_frame(720, 1280, 40 + i) for i in range(B)])
    ex_g, ex_c = orbfe.ORBextractor(2000), O.Extractor(2000)
    ex_g(frames[0], None, (0, 1000))
    ex_g.set_max_bytes(int(ex_g.frame_geometry()["per_frame_bytes"]) * 4)
    n, mono, kps, desc = ex_g.extract_batch(frames, (0, 1000))
    for i in range(B):
        mo, ko, do = ex_c(frames[i], (0, 1000))
        assert mono[i] == mo and n[i] == len(ko) and 0 < mo < len(ko)       # lapping {0,1000} splits 1280-px frames
        assert kps[i, :n[i]].tobytes() == ko.tobytes() and np.array_equal(desc[i, :n[i]], do)


@pytest.mark.parametrize("nf", [5000, 10000, 30000])
def test_large_nfeatures_octree_tables(orbfe, nf):
    """mpIniORBextractor uses 5 x nFeatures (Tracking.cc:665).  10000 features still fit the shared-memory
    quadtree tables (> 48 KB opt-in), 30000 go through the global-memory tables."""
    img = synth.noise_frame(480, 752, 12)
    _check_frame(orbfe.ORBextractor(nf), O.Extractor(nf), img, (0, 1000), stages=False)


def test_unusual_thresholds_and_full_hd(orbfe):
    """minThFAST above iniThFAST, a zero threshold (cv::FAST then needs response > 0), and a 1920x1080 frame."""
    img = synth.synth_frame(480, 640, 31)
    for (ini, mn) in [(7, 20), (20, 0), (0, 0), (255, 1), (12, 12), (254, 127)]:
        _check_frame(orbfe.ORBextractor(800, 1.2, 8, ini, mn), O.Extractor(800, 1.2, 8, ini, mn), img, (0, 0))
    # thresholds on both sides of 128 (the dense reject's byte compare handles the top bit separately) on a
    # high-contrast frame, where such corners exist
    hc = np.where(synth.noise_frame(480, 640, 33) > 127, 255, 0).astype(np.uint8)
    for (ini, mn) in [(140, 126), (200, 128), (127, 100)]:
        _check_frame(orbfe.ORBextractor(800, 1.2, 8, ini, mn), O.Extractor(800, 1.2, 8, ini, mn), hc, (0, 0))
    big = synth.synth_frame(1080, 1920, 32)
    _check_frame(orbfe.ORBextractor(4000), O.Extractor(4000), big, (0, 1000), stages=False)


def test_small_host_batches_graph_replay_and_chunking(orbfe):
    """Host-pointer calls of 1-4 frames replay a captured CUDA graph when they fit one chunk; changing the lapping area,
    the batch size or the image size re-captures; a 4-frame call that needs several chunks takes the chunked path."""
    ex_g, ex_c = orbfe.ORBextractor(1000), O.Extractor(1000)
    frames = np.stack([synth.synth_frame(480, 752, 60 + i) for i in range(4)])

    def check(fr, lap):
        n, mono, kps, desc = ex_g.extract_batch(fr, lap)
        for i in range(len(fr)):
            mo, ko, do = ex_c(fr[i], lap)
            assert mono[i] == mo and n[i] == len(ko) and kps[i, :n[i]].tobytes() == ko.tobytes()
            assert np.array_equal(desc[i, :n[i]], do)
    for lap in ((0, 1000), (0, 1000), (0, 0), (100, 411)):      # second call replays, the others re-capture
        check(frames[:3], lap)
    check(frames[:1], (0, 0))
    check(frames, (0, 1000))
    big = np.stack([synth.synth_frame(720, 1280, 70 + i) for i in range(4)])
    ex_g.set_max_bytes(64 << 20)                                 # 22 MB per 1280x720 frame: chunks of 2 -> 1, 1, 2 frames
    check(big, (0, 1000))
    ex_g.set_max_bytes(6 << 30)
    check(big[:2], (0, 1000))                                    # back to one chunk: graph again, new geometry
    check(frames[:2], (0, 1000))


@pytest.mark.parametrize("per", [2, 3, 16, 1000])
def test_fast_cells_per_warp(orbfe, per, monkeypatch):
    """k_fast_cells walks several cells per warp, one TMA box per cell on the warp's own mbarrier (the phase parity
    alternates per cell); large batches pick 2-8 cells per warp on their own, here the count is forced on single
    frames, including odd counts, a count that crosses pyramid levels and a handful of warps for the whole frame."""
    monkeypatch.setenv("ORBFE_FAST_CELLS_PER_WARP", str(per))
    for (h, w, seed) in ((480, 752, 31), (241, 323, 32)):
        img = synth.synth_frame(h, w, seed)
        _check_frame(orbfe.ORBextractor(1000), O.Extractor(1000), img, (0, 1000), stages=True)


def test_streaming_submit_wait(orbfe):
    """orbfe_extract_batch_submit / _wait: up to four host batches in flight, waited for in FIFO order, each bit-exact
    with the oracle; chunked batches (several chunks per submit, slots alternating across submits), single-chunk
    batches, a size change in mid-stream (the handle drains and re-lays out), and the error paths."""
    import torch
    from orbfe._lib import KP_DTYPE
    ex_g, ex_c = orbfe.ORBextractor(1000), O.Extractor(1000)
    cap = ex_g.capacity
    lap = (0, 1000)

    def pinned_out(B):
        return (torch.empty(B, dtype=torch.int32).pin_memory().numpy(), torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory().numpy().view(KP_DTYPE).reshape(B, cap),
                torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory().numpy())

    def check(frames, out):
        n, mono, kps, desc = out
        for i in range(len(frames)):
            mo, ko, do = ex_c(frames[i], lap)
            assert mono[i] == mo and n[i] == len(ko) and kps[i, :n[i]].tobytes() == ko.tobytes()
            assert np.array_equal(desc[i, :n[i]], do)

    with pytest.raises(Exception):
        ex_g.extract_batch_wait()                                  # nothing in flight
    batches = [torch.from_numpy(np.stack([synth.synth_frame(480, 752, 300 + 10 * k + i) for i in range(5)])).pin_memory()
               for k in range(4)]
    outs = [pinned_out(5) for _ in range(4)]
    for chunk_bytes in (64 << 20, 6 << 30):                        # 64 MB: chunks of 1-2 frames; 6 GB: one chunk per batch
        ex_g.set_max_bytes(chunk_bytes)
        for o in outs:
            for a in o:
                a.view(np.uint8)[...] = 0xEE
        for k in range(4):
            ex_g.extract_batch_submit(batches[k], lap, outs[k])
        with pytest.raises(Exception):
            ex_g.extract_batch_submit(batches[0], lap, pinned_out(5))   # a fifth batch in flight
        for k in range(4):
            ex_g.extract_batch_wait()
            check(batches[k].numpy(), outs[k])
    # steady state as the bench drives it: submit k+1, then wait for k
    ex_g.extract_batch_submit(batches[0], lap, outs[0])
    for k in range(1, 4):
        ex_g.extract_batch_submit(batches[k], lap, outs[k])
        ex_g.extract_batch_wait()
    ex_g.extract_batch_wait()
    for k in range(4):
        check(batches[k].numpy(), outs[k])
    # another image size while a batch is in flight, then a synchronous call that drains everything
    small = torch.from_numpy(np.stack([synth.synth_frame(240, 320, 400 + i) for i in range(3)])).pin_memory()
    so = pinned_out(3)
    ex_g.extract_batch_submit(batches[1], lap, outs[1])
    ex_g.extract_batch_submit(small, lap, so)
    res = ex_g.extract_batch(batches[2].numpy()[:2], lap)
    check(batches[1].numpy(), outs[1])
    check(small.numpy(), so)
    check(batches[2].numpy()[:2], res)
    with pytest.raises(Exception):
        ex_g.extract_batch_wait()
