"""Product logic on the CPU: the per-point / per-match arithmetic of csrc/kb8.cu (csrc/kb8_core.h) compiled for the
host, against the oracle's restatement of KannalaBrandt8::project / unproject / TriangulateMatches."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from test_oracle_kb8 import P1, P2, _points, _rig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "kb8_core_host.cpp")
SO = os.path.join(HERE, "native", "libkb8_core_host.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def core():
    hdr = os.path.join(HERE, "..", "orb-slam3_byzyh_b200", "csrc", "kb8_core.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", SRC, "-o", SO])
    return C.CDLL(SO)


def _ulps(a, b):
    """Distance in units in the last place between two float32 arrays (same sign assumed where it matters)."""
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def test_project_and_unproject(core):
    rng = np.random.default_rng(0)
    p3 = _points(rng, 20000)
    uv = (rng.random((20000, 2)) * 512).astype(np.float32)
    uv[0] = P1[2:4]
    uv[1] = (5000, -3000)
    for P in (P1, P2):
        a, b = O.kb8_project(P, p3), np.empty((len(p3), 2), np.float32)
        core.kb8_core_project(_p(P), _p(p3), len(p3), _p(b))
        # the oracle calls glibc's atan2f (not correctly rounded: up to 1 ulp off), the product rounds a double atan2
        # once; one ulp of theta or psi moves a pixel coordinate of a few hundred by at most a few 1e-4
        assert (a == b).mean() > 0.8 and np.abs(a - b).max() < 5e-4
        for prec in (1e-6, 1e-3):
            a, b = O.kb8_unproject(P, uv, prec), np.empty((len(uv), 3), np.float32)
            core.kb8_core_unproject(_p(P), C.c_float(prec), _p(uv), len(uv), _p(b))
            assert (a == b).mean() > 0.97 and _ulps(a, b).max() <= 4     # tanf against a rounded double tan


def test_triangulate_matches(core):
    rng = np.random.default_rng(5)
    R12, t12, X1, X2 = _rig(rng, 4000)
    pt1, pt2 = O.kb8_project(P1, X1), O.kb8_project(P2, X2)
    pt2[::3] += rng.normal(scale=1.5, size=pt2[::3].shape).astype(np.float32)     # noisy matches: codes -4 / -5 occur
    pt2[5::50] += 40                                                               # gross outliers
    sig = rng.choice(np.float32([1.0, 1.44, 2.0736, 2.985984]), len(X1))
    unc = rng.choice(np.float32([1.0, 1.44, 2.0736, 2.985984]), len(X1))
    d_o, p_o = O.kb8_triangulate(P1, P2, R12, t12, pt1, pt2, sig, unc)
    d_c, p_c = np.empty_like(d_o), np.empty_like(p_o)
    core.kb8_core_triangulate(_p(P1), C.c_float(1e-6), _p(P2), C.c_float(1e-6), _p(R12), _p(t12), _p(pt1), _p(pt2), _p(sig),
                              _p(unc), len(X1), _p(d_c), _p(p_c))
    assert len(set(np.unique(d_o[d_o < 0]).tolist())) >= 3                         # several rejection codes exercised
    same = (d_o < 0) == (d_c < 0)
    assert same.mean() > 0.999                                                     # a threshold may flip on a 1-ulp ray
    neg = same & (d_o < 0)
    assert np.array_equal(d_o[neg], d_c[neg])
    pos = same & (d_o > 0)
    assert np.allclose(d_o[pos], d_c[pos], rtol=1e-4, atol=1e-5) and np.allclose(p_o[pos], p_c[pos], rtol=1e-4, atol=1e-5)
    assert np.isnan(p_c[d_c < 0]).all()
