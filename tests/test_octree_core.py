"""Product logic on the CPU: the CTA-parallel DistributeOctTree (csrc/octree_core.h) compiled
for the host as a single thread must reproduce the oracle (= verbatim reference) retained set
AND list order; its std::sort emulation must equal the real libstdc++ std::sort."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import synth
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "octree_core_host.cpp")
SO = os.path.join(HERE, "native", "liboctree_core_host.so")


@pytest.fixture(scope="module")
def core():
    hdr = os.path.join(HERE, "..", "orb-slam3_byzyh_b200", "csrc", "octree_core.h")
    if (not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", SRC, "-o", SO])
    return C.CDLL(SO)


def _run(core, xys, minX, maxX, minY, maxY, N):
    xys = np.ascontiguousarray(xys, np.int32)
    keep = np.empty(N + 64, np.int32)
    n = core.octree_core_host(xys.ctypes.data_as(C.c_void_p), len(xys), minX, maxX, minY, maxY, N,
                              keep.ctypes.data_as(C.c_void_p), len(keep))
    return keep[:n].copy()


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 18, 33, 100, 257, 1000, 5000])
def test_std_sort_emulation(core, n):
    rng = np.random.default_rng(n)
    for trial in range(20):
        hi = int(rng.choice([2, 4, 16, 1000]))
        keys = rng.integers(0, hi, n).astype(np.uint64)
        a = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        b = a.copy()
        core.octree_core_sort(a.ctypes.data_as(C.c_void_p), n)
        core.octree_core_sort_ref(b.ctypes.data_as(C.c_void_p), n)
        assert np.array_equal(a, b)
        e = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        out = np.zeros(max(n, 1), np.uint64)
        core.octree_core_sort_cta(e.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n)
        assert np.array_equal(out[:n], b)
        c, d = b[::-1].copy(), b[::-1].copy()
        core.octree_core_heapsort(c.ctypes.data_as(C.c_void_p), n)
        core.octree_core_heapsort_ref(d.ctypes.data_as(C.c_void_p), n)
        assert np.array_equal(c, d)


def test_std_sort_depth_limit_path(core):
    # "median-of-3 killer" style input forces the heapsort fallback in introsort
    n = 4096
    half = n // 2
    keys = np.zeros(n, np.uint64)
    for i in range(half):
        if i % 2 == 0:
            keys[i] = i + 1
        else:
            keys[i] = half + i + (half % 2)
        keys[half + i] = 2 * (i + 1)
    a = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    b = a.copy()
    core.octree_core_sort(a.ctypes.data_as(C.c_void_p), n)
    core.octree_core_sort_ref(b.ctypes.data_as(C.c_void_p), n)
    assert np.array_equal(a, b)
    e = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    out = np.zeros(n, np.uint64)
    core.octree_core_sort_cta(e.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n)
    assert np.array_equal(out, b)


@pytest.mark.parametrize("seed", range(40))
def test_random_candidates(core, seed):
    rng = np.random.default_rng(seed)
    W, H = int(rng.integers(60, 1300)), int(rng.integers(60, 700))
    if W < H // 2 + 1:      # nIni == 0 is UB in the reference
        W = H
    n = int(rng.integers(1, min(8000, W * H // 4)))
    pos = rng.choice(W * H, size=n, replace=False)
    xys = np.stack([pos % W, pos // W, rng.integers(7, 40, n)], 1).astype(np.int32)
    xys = xys[np.lexsort((xys[:, 0], xys[:, 1]))]
    N = int(rng.integers(1, 600))
    got = _run(core, xys, 16, 16 + W, 16, 16 + H, N)
    exp = O.octree(xys, 16, 16 + W, 16, 16 + H, N)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("h,w,nf", [(480, 752, 1000), (512, 512, 1500), (720, 1280, 2000), (480, 752, 5000)])
def test_real_fast_candidates(core, h, w, nf):
    img = synth.synth_frame(h, w, 5)
    ex = O.Extractor(nf)
    ex(img, (0, 0))
    nfl = ex.tables()["nfeatures"]
    for lvl in range(8):
        L = ex.level(lvl)
        got = _run(core, L["cands"], 16, L["w"] - 16, 16, L["h"] - 16, int(nfl[lvl]))
        kp = L["kps"]
        exp = np.stack([kp["x"] - 16, kp["y"] - 16, kp["response"]], 1).astype(np.int32)
        assert np.array_equal(L["cands"][got], exp)
