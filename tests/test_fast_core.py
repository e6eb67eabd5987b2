"""Product logic on the CPU: the packed two-pixels-per-register FAST-9/16 score network of
csrc/fast_core.h (host lane emulation) against the oracle's plain FAST restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "fast_core_host.cpp")
SO = os.path.join(HERE, "native", "libfast_core_host.so")


@pytest.fixture(scope="module")
def core():
    hdr = os.path.join(HERE, "..", "orb-slam3_byzyh_b200", "csrc", "fast_core.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", SRC, "-o", SO])
    return C.CDLL(SO)


@pytest.mark.parametrize("kind", ["noise", "lowcontrast", "spikes", "flat"])
@pytest.mark.parametrize("sub", [0, 7, 20])
def test_margin_map_equals_oracle(core, kind, sub):
    rng = np.random.default_rng(hash(kind) % 1000 + sub)
    h, w = 48, 62
    if kind == "noise":
        img = rng.integers(0, 256, (h, w))
    elif kind == "lowcontrast":
        img = 120 + rng.integers(0, 14, (h, w))
    elif kind == "spikes":
        img = np.where(rng.uniform(size=(h, w)) < 0.12, 200 + rng.integers(0, 56, (h, w)), 30 + rng.integers(0, 9, (h, w)))
    else:
        img = np.full((h, w), 255)
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((h, w), np.uint8)
    core.fast_core_margins(img.ctypes.data_as(C.c_void_p), w, h, sub, out.ctypes.data_as(C.c_void_p))
    ref = np.zeros((h, w), np.int64)
    c = O.fast(img, 0, nms=False)            # every pixel with best > 0, response = best-1
    ref[c[:, 1], c[:, 0]] = c[:, 2] + 1
    exp = np.maximum(ref - sub, 0)
    assert np.array_equal(out[3:-3, 3:-3].astype(np.int64), exp[3:-3, 3:-3])
    # the raw-value formulation used by k_fast_score (no per-ring differences)
    out2 = np.zeros((h, w), np.uint8)
    core.fast_core_margins_raw(img.ctypes.data_as(C.c_void_p), w, h, sub, out2.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out2[3:-3, 3:-3].astype(np.int64), exp[3:-3, 3:-3])
    # the pair-sharing formulation (arcs k, k+1 share eight ring positions)
    out3 = np.zeros((h, w), np.uint8)
    core.fast_core_margins_pair(img.ctypes.data_as(C.c_void_p), w, h, sub, out3.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out3[3:-3, 3:-3].astype(np.int64), exp[3:-3, 3:-3])
    out4 = np.zeros((h, w), np.uint8)
    core.fast_core_margins_pair_raw(img.ctypes.data_as(C.c_void_p), w, h, sub, out4.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out4[3:-3, 3:-3].astype(np.int64), exp[3:-3, 3:-3])
    out5 = np.zeros((h, w), np.uint8)   # the biased formulation k_fast_cells uses (no packed subtraction)
    core.fast_core_margins_pair_raw_biased(img.ctypes.data_as(C.c_void_p), w, h, sub, out5.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out5[3:-3, 3:-3].astype(np.int64), exp[3:-3, 3:-3])


@pytest.mark.parametrize("t", [0, 1, 7, 20, 126, 127, 128, 200, 254])
def test_compass_reject_is_exact_and_necessary(core, t):
    """fc_compass4 (the dense early reject of k_fast_cells): equals the plain statement of the test for every byte
    value and threshold, and never rejects a pixel that cv::FAST calls a corner at that threshold."""
    rng = np.random.default_rng(t)
    h, w = 40, 3 + 4 * 14 + 3
    imgs = [rng.integers(0, 256, (h, w)), np.where(rng.uniform(size=(h, w)) < 0.5, 255, 0),
            128 + rng.integers(-2, 3, (h, w)) * (t + 1) // 2]
    for img in imgs:
        img = np.ascontiguousarray(np.clip(img, 0, 255), np.uint8)
        out = np.zeros((h, w), np.uint8)
        core.fast_core_compass(img.ctypes.data_as(C.c_void_p), w, h, t, out.ctypes.data_as(C.c_void_p))
        I = img.astype(np.int64)
        c = I[3:-3, 3:-3]
        d = lambda dy, dx: np.abs(I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] - c) > t
        exp = (d(-3, 0) | d(3, 0)) & (d(0, -3) | d(0, 3))
        assert np.array_equal(out[3:-3, 3:-3].astype(bool), exp)
        corners = O.fast(img, t, nms=False)
        assert out[corners[:, 1], corners[:, 0]].all()
