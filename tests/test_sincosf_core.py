"""Product logic on the CPU: csrc/sincosf_core.h (the sinf / cosf restatement k_describe rotates the BRIEF pattern with,
ORBextractor.cc:157) compiled for the host, against the libm the reference's own build calls."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "sincosf_core_host.cpp")
SO = os.path.join(HERE, "native", "libsincosf_core_host.so")


@pytest.fixture(scope="module")
def core():
    hdr = os.path.join(HERE, "..", "orb-slam3_byzyh_b200", "csrc", "sincosf_core.h")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", SRC, "-o", SO])
    return C.CDLL(SO)


def _libm(y):
    libm = C.CDLL("libm.so.6")
    libm.sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    sn, cs = np.empty(len(y), np.float32), np.empty(len(y), np.float32)
    s, c = C.c_float(), C.c_float()
    for i, v in enumerate(y):
        libm.sincosf(float(v), C.byref(s), C.byref(c))
        sn[i], cs[i] = s.value, c.value
    return sn, cs


def _ours(core, y):
    sn, cs = np.empty(len(y), np.float32), np.empty(len(y), np.float32)
    core.host_sincosf(y.ctypes.data_as(C.c_void_p), len(y), sn.ctypes.data_as(C.c_void_p), cs.ctypes.data_as(C.c_void_p))
    return sn, cs


def test_matches_libm_bit_for_bit(core):
    rng = np.random.default_rng(0)
    factor = np.float32(np.pi / 180.0)
    deg = np.concatenate([rng.uniform(0, 360, 300000).astype(np.float32),            # what IC_Angle / fastAtan2 hands over
                          np.arange(0, 360, 0.25, dtype=np.float32),
                          np.float32([0, 1e-6, 1e-3, 0.2, 42.9, 43.0, 44.99, 45, 89.99, 90, 179.99, 180, 193.49879455566406,
                                      270, 359.99])])
    y = np.ascontiguousarray(deg * factor, np.float32)
    y = np.concatenate([y, rng.uniform(-100, 100, 100000).astype(np.float32), np.float32([2 ** -13, 2 ** -12, 0.7499, 0.75, 119.9])])
    sn, cs = _ours(core, y)
    ls, lc = _libm(y)
    assert np.array_equal(sn.view(np.uint32), ls.view(np.uint32))
    assert np.array_equal(cs.view(np.uint32), lc.view(np.uint32))


def test_libm_is_not_correctly_rounded_there(core):
    """Why the restatement exists: on the angle the random parity sweep found, libm's sinf is 0.517 ulp off, the rounded
    fp64 value differs by one ulp, and that moves a rotated BRIEF sample from row 3 to row 4."""
    y = np.float32([np.float32(193.49879455566406) * np.float32(np.pi / 180.0)])
    sn, cs = _ours(core, y)
    assert sn[0] != np.float32(np.sin(np.float64(y[0]))) and cs[0] == np.float32(np.cos(np.float64(y[0])))
    row = np.float32(np.float32(10) * sn[0]) + np.float32(np.float32(-6) * cs[0])
    assert np.rint(np.float32(row)) == 3
