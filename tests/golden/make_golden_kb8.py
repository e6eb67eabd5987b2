#!/usr/bin/env python
"""Generates tests/golden/kb8_ref.npz from the REFERENCE ITSELF: KannalaBrandt8::project (cv::Point3f overload) and
::unproject of /root/reference/src/CameraModels/KannalaBrandt8.cpp, compiled verbatim into oracle/_ref/libref_kb8.so
(oracle/ref_build.sh).  Seeded inputs and the reference outputs are stored.  Run in the build container:
    python tests/golden/make_golden_kb8.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from test_oracle_kb8 import P1, P2, _points  # noqa: E402


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_kb8.so"))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(2024)
    p3 = _points(rng, 3000)
    uv = (rng.random((3000, 2)) * 512).astype(np.float32)
    uv[0] = P1[2:4]
    uv[1] = (5000, -3000)
    out = {"p3d": p3, "uv": uv, "P1": P1, "P2": P2}
    for name, P in (("1", P1), ("2", P2)):
        proj = np.empty((len(p3), 2), np.float32)
        ref.ref_kb8_project(p(P), p(p3), len(p3), p(proj))
        rays = np.empty((len(uv), 3), np.float32)
        ref.ref_kb8_unproject(p(P), C.c_float(1e-6), p(uv), len(uv), p(rays))
        out["project" + name], out["unproject" + name] = proj, rays
    np.savez_compressed(os.path.join(HERE, "kb8_ref.npz"), **out)
    print("wrote kb8_ref.npz,", os.path.getsize(os.path.join(HERE, "kb8_ref.npz")), "bytes")


if __name__ == "__main__":
    main()
