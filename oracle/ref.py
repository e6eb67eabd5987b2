"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes front for oracle/_ref/libref_orbextractor.so: the reference's own
src/ORBextractor.cc compiled verbatim against oracle/cvshim (see oracle/ref_build.sh).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import KP_DTYPE, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_orbextractor.so")
_LIB = None


def available():
    if not os.path.exists(_SO) and os.path.exists("/root/reference/src/ORBextractor.cc"):
        subprocess.call(["sh", os.path.join(_HERE, "ref_build.sh")])
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_orbextractor.so not built (needs /root/reference)")
        _LIB = C.CDLL(_SO)
        _LIB.ref_extractor_create.restype = C.c_void_p
        _LIB.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.ref_extractor_destroy.argtypes = [C.c_void_p]
    return _LIB


class RefExtractor:
    """The reference's ORB_SLAM3::ORBextractor itself."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = lib().ref_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_extractor_destroy(self.h)
            self.h = None

    def __call__(self, image, lapping=(0, 0)):
        image = np.ascontiguousarray(image, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rows, cols = image.shape if image.size else (0, 0)
        mono = lib().ref_extract(C.c_void_p(self.h), _p(image), rows, cols,
                                 image.strides[0] if image.size else 0, lapping[0], lapping[1],
                                 _p(kps), _p(desc), cap, C.byref(n))
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level_padded(self, lvl):
        w, h = C.c_int(), C.c_int()
        lib().ref_level_dims(C.c_void_p(self.h), lvl, C.byref(w), C.byref(h))
        out = np.empty((h.value + 38, w.value + 38), np.uint8)
        lib().ref_level_padded(C.c_void_p(self.h), lvl, _p(out))
        return out

    def octree(self, xys, minX, maxX, minY, maxY, N):
        xys = np.ascontiguousarray(xys, np.int32)
        out = np.empty((max(len(xys), 8), 3), np.int32)
        n = lib().ref_octree(C.c_void_p(self.h), _p(xys), len(xys), minX, maxX, minY, maxY, N, _p(out), len(out))
        return out[:n].copy()
