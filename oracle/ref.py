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


# ------------------------------------------------------------------------------------------------
# oracle/_ref/libref_orbmatcher.so: the reference's own matcher functions (function bodies cut out of
# src/ORBmatcher.cc, src/Frame.cc, src/MapPoint.cc at build time and compiled verbatim against
# oracle/refshim; see oracle/ref_build.sh, oracle/ref_slices.py, oracle/ref_match_driver.cpp).
_MSO = os.path.join(_HERE, "_ref", "libref_orbmatcher.so")
_MLIB = None


def matcher_available():
    if not os.path.exists(_MSO) and os.path.exists("/root/reference/src/ORBmatcher.cc"):
        subprocess.call(["sh", os.path.join(_HERE, "ref_build.sh")])
    return os.path.exists(_MSO)


def mlib():
    global _MLIB
    if _MLIB is None:
        if not matcher_available():
            raise RuntimeError("oracle/_ref/libref_orbmatcher.so not built (needs /root/reference)")
        _MLIB = C.CDLL(_MSO)
        _MLIB.refm_frame_create.restype = C.c_void_p
    return _MLIB


def _f(x):
    return C.c_float(float(x))


def _opt(a, dt):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dt)
    return a, _p(a)


def set_bounds(bounds):
    """bounds = (minX, minY, maxX, maxY): Frame's static image bounds and grid cell inverses (Frame.cc:303-305)."""
    b = [np.float32(v) for v in bounds]
    gw = np.float32(64) / np.float32(b[2] - b[0])
    gh = np.float32(48) / np.float32(b[3] - b[1])
    mlib().refm_set_bounds(_f(b[0]), _f(b[1]), _f(b[2]), _f(b[3]), _f(gw), _f(gh))


class RefFrame:
    """A reference Frame (shim class, reference member functions) built from plain arrays."""

    def __init__(self, keys, desc, scale_factors, uright=None, right=None, mb=0.0, mbf=0.0, scale=1.2):
        """right = (keysR, descR, l2r, r2l) makes it a fisheye stereo frame (Nleft != -1)."""
        keys = np.ascontiguousarray(keys)
        desc = np.ascontiguousarray(desc, np.uint8)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        ur, pur = _opt(uright, np.float32)
        self.n = len(keys)
        if right is not None:
            kr = np.ascontiguousarray(right[0]); dr = np.ascontiguousarray(right[1], np.uint8)
            l2r = np.ascontiguousarray(right[2], np.int32); r2l = np.ascontiguousarray(right[3], np.int32)
            args = (_p(kr), len(kr), _p(dr), _p(l2r), _p(r2l))
            self.n += len(kr)
        else:
            args = (None, -1, None, None, None)
        self.h = mlib().refm_frame_create(_p(keys), len(keys), _p(desc), pur, *args, _p(sf), len(sf),
                                          _f(np.log(np.float32(scale))), _f(mb), _f(mbf))

    def __del__(self):
        if getattr(self, "h", None):
            mlib().refm_frame_destroy(C.c_void_p(self.h))
            self.h = None

    def set_pose(self, t):
        mlib().refm_frame_pose(C.c_void_p(self.h), _f(t[0]), _f(t[1]), _f(t[2]))

    def set_trl(self, t):
        mlib().refm_frame_trl(C.c_void_p(self.h), _f(t[0]), _f(t[1]), _f(t[2]))

    def set_mappoints(self, has, nobs=None, xyz=None, desc=None, outlier=None, bad=None, min_dist=None, max_dist=None):
        keep = [_opt(has, np.uint8), _opt(nobs, np.int32), _opt(xyz, np.float32), _opt(desc, np.uint8),
                _opt(outlier, np.uint8), _opt(bad, np.uint8), _opt(min_dist, np.float32), _opt(max_dist, np.float32)]
        assert len(keep[0][0]) == self.n
        mlib().refm_frame_mappoints(C.c_void_p(self.h), *[k[1] for k in keep])

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1, right=False):
        out = np.empty(self.n + 1, np.int32)
        n = mlib().refm_features_in_area(C.c_void_p(self.h), _f(x), _f(y), _f(r), int(min_level), int(max_level),
                                         int(right), _p(out), len(out))
        return out[:n].copy()

    def predict_scale(self, max_distance, dist):
        return mlib().refm_predict_scale(C.c_void_p(self.h), _f(max_distance), _f(dist))

    def search_mappoints(self, mp, th, far_points, th_far, nnratio):
        """mp: dict with in_view, in_view_r, depth, bad, nobs, proj_x, proj_y, proj_xr, proj_yr, level, level_r,
        view_cos, view_cos_r, desc (missing right-camera entries = None)."""
        order = [("in_view", np.uint8), ("in_view_r", np.uint8), ("depth", np.float32), ("bad", np.uint8),
                 ("nobs", np.int32), ("proj_x", np.float32), ("proj_y", np.float32), ("proj_xr", np.float32),
                 ("proj_yr", np.float32), ("level", np.int32), ("level_r", np.int32), ("view_cos", np.float32),
                 ("view_cos_r", np.float32), ("desc", np.uint8)]
        keep = [_opt(mp.get(k), dt) for k, dt in order]
        slots = np.empty(self.n, np.int32)
        n = mlib().refm_search_mappoints(C.c_void_p(self.h), len(keep[0][0]), *[k[1] for k in keep], _f(th),
                                         int(far_points), _f(th_far), _f(nnratio), _p(slots))
        return n, slots

    def search_lastframe(self, last, th, mono, nnratio, check_ori):
        slots = np.empty(self.n, np.int32)
        n = mlib().refm_search_lastframe(C.c_void_p(self.h), C.c_void_p(last.h), _f(th), int(mono), _f(nnratio),
                                         int(check_ori), _p(slots))
        return n, slots

    def search_keyframe(self, kf, already_found, th, orb_dist, nnratio, check_ori):
        af, paf = _opt(already_found, np.uint8)
        slots = np.empty(self.n, np.int32)
        n = mlib().refm_search_keyframe(C.c_void_p(self.h), C.c_void_p(kf.h), paf, _f(th), int(orb_dist),
                                        _f(nnratio), int(check_ori), _p(slots))
        return n, slots


def search_for_initialization(f1, f2, prev_matched, window_size, nnratio, check_ori):
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.empty(len(prev), np.int32)
    n = mlib().refm_search_init(C.c_void_p(f1.h), C.c_void_p(f2.h), _p(prev), _p(m12), int(window_size),
                                _f(nnratio), int(check_ori))
    return n, m12, prev


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return mlib().refm_descriptor_distance(_p(a), _p(b))


def stereo(img_l, img_r, mbf, mb, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
    """Reference ORBextractor on both images + Frame::ComputeStereoMatches.
    Returns keysL, descL, keysR, descR, uRight, depth."""
    img_l = np.ascontiguousarray(img_l, np.uint8); img_r = np.ascontiguousarray(img_r, np.uint8)
    cap = nfeatures + 64 * nlevels + 64
    kl, kr = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dl, dr = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    ur, dp = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    nr = C.c_int(0)
    n = mlib().refm_stereo(_p(img_l), _p(img_r), img_l.shape[0], img_l.shape[1], nfeatures, _f(scale), nlevels, ini, mn,
                           _f(mbf), _f(mb), _p(kl), _p(dl), _p(kr), _p(dr), cap, C.byref(nr), _p(ur), _p(dp))
    assert n >= 0
    return kl[:n].copy(), dl[:n].copy(), kr[:nr.value].copy(), dr[:nr.value].copy(), ur[:n].copy(), dp[:n].copy()


# ---- keyframe-side searches (Fuse x2, SearchBySim3, Sim3 SearchByProjection) ----------------------------
def _kf_points_args(P, m, with_has=False, with_inkf=False, with_nobs=False):
    """P: dict of per-candidate arrays (has, bad, in_kf, xyz, normal, min_dist, max_dist, nobs, desc)."""
    keep = []

    def a(name, dt):
        arr, p = _opt(P.get(name), dt)
        keep.append(arr)
        return p
    args = []
    if with_has:
        args.append(a("has", np.uint8))
    args.append(a("bad", np.uint8))
    if with_inkf:
        args.append(a("in_kf", np.uint8))
    args += [a("xyz", np.float32), a("normal", np.float32), a("min_dist", np.float32), a("max_dist", np.float32)]
    if with_nobs:
        args.append(a("nobs", np.int32))
    args.append(a("desc", np.uint8))
    return args, keep


def set_camera(frame, fx, fy, cx, cy):
    mlib().refm_frame_camera(C.c_void_p(frame.h), _f(fx), _f(fy), _f(cx), _f(cy))


def kf_predict_scale(frame, max_distance, dist):
    return mlib().refm_kf_predict_scale(C.c_void_p(frame.h), _f(max_distance), _f(dist))


def fuse(kf, P, th):
    m = len(P["xyz"])
    args, keep = _kf_points_args(P, m, with_has=True, with_inkf=True, with_nobs=True)
    cap = 4 * m + 16
    actions = np.zeros((cap, 3), np.int32)
    na = C.c_int(0)
    n = mlib().refm_fuse(C.c_void_p(kf.h), m, *args, _f(th), _p(actions), cap, C.byref(na))
    return n, actions[:na.value].copy()


def fuse_sim3(kf, scw, P, th):
    m = len(P["xyz"])
    args, keep = _kf_points_args(P, m)
    cap = 4 * m + 16
    actions = np.zeros((cap, 3), np.int32)
    na = C.c_int(0)
    repl = np.empty(m, np.int32)
    n = mlib().refm_fuse_sim3(C.c_void_p(kf.h), _f(scw[0]), _f(scw[1]), _f(scw[2]), _f(scw[3]), m, *args, _f(th),
                              _p(repl), _p(actions), cap, C.byref(na))
    return n, repl, actions[:na.value].copy()


def search_kf_sim3(kf, scw, P, matched, th, ratio_hamming):
    m = len(P["xyz"])
    args, keep = _kf_points_args(P, m)
    matched = np.ascontiguousarray(matched, np.uint8)
    slots = np.empty(len(matched), np.int32)
    n = mlib().refm_search_kf_sim3(C.c_void_p(kf.h), _f(scw[0]), _f(scw[1]), _f(scw[2]), _f(scw[3]), m, *args,
                                   _p(matched), int(th), _f(ratio_hamming), _p(slots))
    return n, slots


def search_by_sim3(kf1, kf2, s12, pre12, th):
    pre12 = np.ascontiguousarray(pre12, np.int32)
    out = np.empty(len(pre12), np.int32)
    n = mlib().refm_search_by_sim3(C.c_void_p(kf1.h), C.c_void_p(kf2.h), _f(s12[0]), _f(s12[1]), _f(s12[2]), _f(s12[3]),
                                   _p(pre12), _f(th), _p(out))
    return n, out


# ---- DBoW2 itself (compiled verbatim into libref_orbmatcher.so) and the reference's SearchByBoW ------------
class RefVocabulary:
    def __init__(self, path):
        mlib().refd_voc_load.restype = C.c_void_p
        self.h = mlib().refd_voc_load(path.encode())
        if not self.h:
            raise RuntimeError("loadFromTextFile failed: " + path)

    def __del__(self):
        if getattr(self, "h", None):
            mlib().refd_voc_destroy(C.c_void_p(self.h))
            self.h = None

    def size(self):
        return mlib().refd_voc_size(C.c_void_p(self.h))

    def transform_features(self, desc, levelsup):
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        word, nid, w = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64)
        mlib().refd_voc_transform_features(C.c_void_p(self.h), _p(desc), n, int(levelsup), _p(word), _p(w), _p(nid))
        return word, w, nid

    def transform(self, desc, levelsup):
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        ids, vals = np.empty(n + 1, np.uint32), np.empty(n + 1, np.float64)
        nodes, start, feat = np.empty(n + 1, np.uint32), np.empty(n + 2, np.int32), np.empty(n + 1, np.uint32)
        nw, nn = C.c_int(0), C.c_int(0)
        mlib().refd_voc_transform(C.c_void_p(self.h), _p(desc), n, int(levelsup), C.byref(nw), _p(ids), _p(vals),
                                  C.byref(nn), _p(nodes), _p(start), _p(feat))
        nw, nn = nw.value, nn.value
        return (ids[:nw].copy(), vals[:nw].copy()), (nodes[:nn].astype(np.int32), start[:nn + 1].copy(),
                                                     feat[:start[nn]].astype(np.int32))


def compute_bow(frame, voc, levelsup=4):
    mlib().refm_frame_compute_bow(C.c_void_p(frame.h), C.c_void_p(voc.h), int(levelsup))


def search_by_bow_kf_f(kf, f, nnratio, check_ori):
    out = np.empty(f.n, np.int32)
    n = mlib().refm_search_by_bow_kf_f(C.c_void_p(kf.h), C.c_void_p(f.h), _f(nnratio), int(check_ori), _p(out))
    return n, out


def search_by_bow_kf_kf(kf1, kf2, nnratio, check_ori):
    out = np.empty(kf1.n, np.int32)
    n = mlib().refm_search_by_bow_kf_kf(C.c_void_p(kf1.h), C.c_void_p(kf2.h), _f(nnratio), int(check_ori), _p(out))
    return n, out


def search_for_triangulation(kf1, kf2, only_stereo, coarse, check_ori):
    """-> nmatches, matches12, F12 (3x3), epipole (2)."""
    out = np.empty(kf1.n, np.int32)
    f12, ep = np.empty(9, np.float32), np.empty(2, np.float32)
    n = mlib().refm_search_for_triangulation(C.c_void_p(kf1.h), C.c_void_p(kf2.h), int(only_stereo), int(coarse), _f(0.6),
                                             int(check_ori), _p(out), _p(f12), _p(ep))
    return n, out, f12, ep


def distinctive(desc, kf_start, rows, kf_bad):
    """MapPoint::ComputeDistinctiveDescriptors per point -> chosen descriptors (n_points x 32; zeros where none)."""
    desc = np.ascontiguousarray(desc, np.uint8)
    kf_start = np.ascontiguousarray(kf_start, np.int32); rows = np.ascontiguousarray(rows, np.int32)
    kf_bad = np.ascontiguousarray(kf_bad, np.uint8)
    out = np.zeros((len(kf_start) - 1, 32), np.uint8)
    mlib().refm_distinctive(_p(desc), _p(kf_start), _p(rows), _p(kf_bad), len(kf_start) - 1, _p(out))
    return out
