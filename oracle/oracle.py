"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front for oracle/liborb_oracle.so, the CPU restatement of the reference's ORB
front-end (see oracle/orb_oracle.h, oracle/match_oracle.h, oracle/cvprims.h for the
reference file:line each function follows).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product path
(orb-slam3_byzyh_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    so = os.path.join(_HERE, "liborb_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liborb_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_extractor_create.restype = C.c_void_p
        _LIB.oracle_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.oracle_extractor_destroy.argtypes = [C.c_void_p]
        _LIB.oracle_ic_angle.restype = C.c_float
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class FrameViewC(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys", C.c_void_p), ("uright", C.c_void_p),
                ("desc", C.c_void_p), ("min_x", C.c_float), ("min_y", C.c_float),
                ("max_x", C.c_float), ("max_y", C.c_float), ("grid_w_inv", C.c_float),
                ("grid_h_inv", C.c_float)]


class ProjPointsC(C.Structure):
    _fields_ = [("m", C.c_int32), ("u", C.c_void_p), ("v", C.c_void_p), ("ur", C.c_void_p),
                ("radius", C.c_void_p), ("min_level", C.c_void_p), ("max_level", C.c_void_p),
                ("angle", C.c_void_p), ("valid", C.c_void_p), ("blocks", C.c_void_p),
                ("desc", C.c_void_p)]


class SearchParamsC(C.Structure):
    _fields_ = [("mode", C.c_int32), ("th_accept", C.c_int32), ("nnratio", C.c_float),
                ("check_orientation", C.c_int32)]


# ---------------------------------------------------------------- primitives
def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def border101(src, pad):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * pad, w + 2 * pad), np.uint8)
    lib().oracle_border101(_p(src), w, h, src.strides[0], _p(dst), dst.strides[0], pad)
    return dst


def fast(img, th, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.empty((cap, 3), np.int32)
    n = lib().oracle_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], th, int(nms), _p(out), cap)
    return out[:n].copy()


def blur7(img):
    img = np.ascontiguousarray(img, np.uint8)
    dst = np.empty_like(img)
    lib().oracle_blur7(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(dst), dst.strides[0])
    return dst


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().oracle_fast_atan2(_p(y), _p(x), _p(out), y.size)
    return out


def hamming(a, b):
    return lib().oracle_descriptor_distance(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)))


def knn2(q, t):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.empty((len(q), 2), np.int32)
    dist = np.empty((len(q), 2), np.int32)
    lib().oracle_knn2(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist))
    return idx, dist


def fisheye_matches(q, t):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.empty((len(q), 2), np.int32)
    dist = np.empty((len(q), 2), np.int32)
    match = np.empty(len(q), np.int32)
    lib().oracle_fisheye_matches(_p(q), len(q), _p(t), len(t), _p(match), _p(idx), _p(dist))
    return match, idx, dist


def octree(xys, minX, maxX, minY, maxY, N):
    xys = np.ascontiguousarray(xys, np.int32)
    keep = np.empty(max(len(xys), 8), np.int32)
    n = lib().oracle_octree(_p(xys), len(xys), minX, maxX, minY, maxY, N, _p(keep), len(keep))
    return keep[:n].copy()


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return float(lib().oracle_ic_angle(_p(img), img.strides[0], int(x), int(y)))


def descriptor(img, x, y, angle):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.empty(32, np.uint8)
    lib().oracle_descriptor(_p(img), img.strides[0], int(x), int(y), C.c_float(angle), _p(d))
    return d


# ---------------------------------------------------------------- extractor
class Extractor:
    """Mirror of ORB_SLAM3::ORBextractor on the CPU oracle."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = lib().oracle_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, inv, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        nf = np.empty(n, np.int32)
        um = np.empty(16, np.int32)
        lib().oracle_tables(C.c_void_p(self.h), _p(sc), _p(inv), _p(s2), _p(is2), _p(nf), _p(um))
        return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, nfeatures=nf, umax=um)

    def __call__(self, image, lapping=(0, 0)):
        image = np.ascontiguousarray(image, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rows, cols = image.shape if image.size else (0, 0)
        mono = lib().oracle_extract(C.c_void_p(self.h), _p(image), rows, cols,
                                    image.strides[0] if image.size else 0, lapping[0], lapping[1],
                                    _p(kps), _p(desc), cap, C.byref(n))
        assert n.value <= cap
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level(self, lvl):
        """Intermediates of the last call for one level."""
        w, h, nc, nk, ncell = (C.c_int() for _ in range(5))
        lib().oracle_level_dims(C.c_void_p(self.h), lvl, C.byref(w), C.byref(h), C.byref(nc),
                                C.byref(nk), C.byref(ncell))
        padded = np.empty((h.value + 38, w.value + 38), np.uint8)
        blurred = np.zeros((h.value, w.value), np.uint8)
        lib().oracle_level_images(C.c_void_p(self.h), lvl, _p(padded), _p(blurred))
        cands = np.empty((nc.value, 3), np.int32)
        retry = np.empty(ncell.value, np.uint8)
        kps = np.empty(nk.value, KP_DTYPE)
        lib().oracle_level_lists(C.c_void_p(self.h), lvl, _p(cands), _p(retry), _p(kps))
        return dict(w=w.value, h=h.value, padded=padded, blurred=blurred, cands=cands,
                    cell_retry=retry, kps=kps)


# ---------------------------------------------------------------- matchers
def make_frame_view(keys, desc, uright, bounds, keep):
    """bounds = (minX, minY, maxX, maxY).  `keep` collects arrays that must stay alive."""
    keys = np.ascontiguousarray(keys)
    desc = np.ascontiguousarray(desc, np.uint8)
    keep += [keys, desc]
    fv = FrameViewC()
    fv.n = len(keys)
    fv.keys = keys.ctypes.data
    fv.desc = desc.ctypes.data
    if uright is not None:
        uright = np.ascontiguousarray(uright, np.float32)
        keep.append(uright)
        fv.uright = uright.ctypes.data
    fv.min_x, fv.min_y, fv.max_x, fv.max_y = [np.float32(b) for b in bounds]
    # reference src/Frame.cc:303-305
    fv.grid_w_inv = np.float32(64) / np.float32(np.float32(bounds[2]) - np.float32(bounds[0]))
    fv.grid_h_inv = np.float32(48) / np.float32(np.float32(bounds[3]) - np.float32(bounds[1]))
    return fv


def make_proj_points(pts, keep):
    """pts: dict of arrays u,v,ur,radius,min_level,max_level,angle,valid,blocks,desc."""
    pp = ProjPointsC()
    m = len(pts["u"])
    pp.m = m
    for name, dt in [("u", np.float32), ("v", np.float32), ("ur", np.float32),
                     ("radius", np.float32), ("min_level", np.int32), ("max_level", np.int32),
                     ("angle", np.float32), ("valid", np.uint8), ("blocks", np.uint8),
                     ("desc", np.uint8)]:
        a = np.ascontiguousarray(pts[name], dt)
        keep.append(a)
        setattr(pp, name, a.ctypes.data)
    return pp


def search_by_projection(keys, desc, uright, bounds, pts, mode, th_accept, nnratio,
                         check_orientation, claimed, assigned, scale_factors):
    keep = []
    fv = make_frame_view(keys, desc, uright, bounds, keep)
    pp = make_proj_points(pts, keep)
    prm = SearchParamsC(mode, th_accept, nnratio, int(check_orientation))
    claimed = np.ascontiguousarray(claimed, np.uint8)
    assigned = np.ascontiguousarray(assigned, np.int32).copy()
    bi = np.empty(pp.m, np.int32)
    bd = np.empty(pp.m, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    n = lib().oracle_search_by_projection(C.byref(fv), C.byref(pp), C.byref(prm), _p(sf), len(sf),
                                          _p(claimed), _p(assigned), _p(bi), _p(bd))
    return n, assigned, bi, bd


def features_in_area(keys, bounds, x, y, r, min_level, max_level):
    keep = []
    fv = make_frame_view(keys, np.zeros((len(keys), 32), np.uint8), None, bounds, keep)
    out = np.empty(len(keys) + 1, np.int32)
    n = lib().oracle_features_in_area(C.byref(fv), C.c_float(x), C.c_float(y), C.c_float(r),
                                      min_level, max_level, _p(out), len(out))
    return out[:n].copy()


def stereo_match(exL, exR, keysL, descL, keysR, descR, mbf, mb):
    keysL = np.ascontiguousarray(keysL); keysR = np.ascontiguousarray(keysR)
    descL = np.ascontiguousarray(descL, np.uint8); descR = np.ascontiguousarray(descR, np.uint8)
    ur = np.empty(len(keysL), np.float32)
    dp = np.empty(len(keysL), np.float32)
    lib().oracle_stereo_match(C.c_void_p(exL.h), C.c_void_p(exR.h), _p(keysL), _p(descL), len(keysL),
                              _p(keysR), _p(descR), len(keysR), C.c_float(mbf), C.c_float(mb),
                              _p(ur), _p(dp))
    return ur, dp


def search_by_projection_fisheye(keysL, descL, keysR, descR, bounds, l2r, r2l, ptsL, ptsR, mode, th_accept, nnratio,
                                 check_orientation, claimed, assigned):
    """ptsL: dict as for search_by_projection (u,v,radius,min_level,max_level,angle,valid,blocks,desc);
    ptsR: dict with u,v,radius,min_level,max_level,valid for the right camera."""
    keep = []
    fl = make_frame_view(keysL, descL, None, bounds, keep)
    fr = make_frame_view(keysR, descR, None, bounds, keep)
    full = dict(ptsL)
    full.setdefault("ur", np.zeros(len(ptsL["u"]), np.float32))
    pl = make_proj_points(full, keep)
    fr_pts = dict(full)
    fr_pts.update(ptsR)
    pr = make_proj_points(fr_pts, keep)
    prm = SearchParamsC(mode, th_accept, nnratio, int(check_orientation))
    l2r = np.ascontiguousarray(l2r, np.int32)
    r2l = np.ascontiguousarray(r2l, np.int32)
    claimed = np.ascontiguousarray(claimed, np.uint8)
    assigned = np.ascontiguousarray(assigned, np.int32).copy()
    bl, br = np.empty(pl.m, np.int32), np.empty(pl.m, np.int32)
    n = lib().oracle_search_by_projection_fisheye(C.byref(fl), C.byref(fr), _p(l2r), _p(r2l), C.byref(pl), C.byref(pr),
                                                  C.byref(prm), _p(claimed), _p(assigned), _p(bl), _p(br))
    return n, assigned, bl, br


def search_for_initialization(keys1, desc1, keys2, desc2, bounds, prev_matched, window_size, nnratio, check_orientation):
    keep = []
    f1 = make_frame_view(keys1, desc1, None, bounds, keep)
    f2 = make_frame_view(keys2, desc2, None, bounds, keep)
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.empty(len(keys1), np.int32)
    lib().oracle_search_for_initialization.argtypes = None
    n = lib().oracle_search_for_initialization(C.byref(f1), C.byref(f2), _p(prev), int(window_size), C.c_float(nnratio),
                                               int(check_orientation), _p(m12))
    return n, m12, prev


def search_window(keys, desc, uright, bounds, pts, th_accept, fuse_gate=False, inv_level_sigma2=None):
    """Stateless keyframe-side window search (Fuse x2, SearchBySim3, Sim3 SearchByProjection inner loop)."""
    keep = []
    fv = make_frame_view(keys, desc, uright, bounds, keep)
    full = dict(pts)
    m = len(pts["u"])
    full.setdefault("ur", np.zeros(m, np.float32))
    full.setdefault("angle", np.zeros(m, np.float32))
    full.setdefault("blocks", np.ones(m, np.uint8))
    pp = make_proj_points(full, keep)
    inv = np.ascontiguousarray(inv_level_sigma2 if inv_level_sigma2 is not None else np.zeros(1), np.float32)
    bi, bd = np.empty(m, np.int32), np.empty(m, np.int32)
    lib().oracle_search_window(C.byref(fv), C.byref(pp), int(th_accept), int(fuse_gate), _p(inv), len(inv), _p(bi), _p(bd))
    return bi, bd


def search_by_sim3(keys1, desc1, keys2, desc2, bounds, pts12, pts21, th_accept=100):
    keep = []
    f1 = make_frame_view(keys1, desc1, None, bounds, keep)
    f2 = make_frame_view(keys2, desc2, None, bounds, keep)
    pps = []
    for pts in (pts12, pts21):
        full = dict(pts)
        m = len(pts["u"])
        for k, dt in (("ur", np.float32), ("angle", np.float32)):
            full.setdefault(k, np.zeros(m, dt))
        full.setdefault("blocks", np.ones(m, np.uint8))
        pps.append(make_proj_points(full, keep))
    m12 = np.empty(len(keys1), np.int32)
    n = lib().oracle_search_by_sim3(C.byref(f1), C.byref(f2), C.byref(pps[0]), C.byref(pps[1]), int(th_accept), _p(m12))
    return n, m12


# ---------------------------------------------------------------- bag of words
class Vocabulary:
    """DBoW2 vocabulary tree restated (oracle/bow_oracle.cpp).  parent/desc/weight indexed by node id (0 = root)."""

    def __init__(self, k, L, parent, desc, weight, scoring=0, weighting=0):
        self.k, self.L = k, L
        self.parent = np.ascontiguousarray(parent, np.int32)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.weight = np.ascontiguousarray(weight, np.float64)
        self.scoring, self.weighting = scoring, weighting
        lib().oracle_voc_create.restype = C.c_void_p
        self.h = lib().oracle_voc_create(k, L, scoring, weighting, len(self.parent), _p(self.parent), _p(self.desc),
                                         _p(self.weight))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_voc_destroy(C.c_void_p(self.h))
            self.h = None

    def transform_features(self, desc, levelsup):
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        word, nid, w = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64)
        lib().oracle_bow_transform_features(C.c_void_p(self.h), _p(desc), n, int(levelsup), _p(word), _p(w), _p(nid))
        return word, w, nid

    def transform(self, desc, levelsup):
        """-> (word ids, values), (node ids, start, feat): BowVector and FeatureVector in map order."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        ids, vals = np.empty(n + 1, np.uint32), np.empty(n + 1, np.float64)
        nodes, start, feat = np.empty(n + 1, np.uint32), np.empty(n + 2, np.int32), np.empty(n + 1, np.uint32)
        nw, nn = C.c_int(0), C.c_int(0)
        lib().oracle_bow_transform(C.c_void_p(self.h), _p(desc), n, int(levelsup), C.byref(nw), _p(ids), _p(vals),
                                   C.byref(nn), _p(nodes), _p(start), _p(feat))
        nw, nn = nw.value, nn.value
        return (ids[:nw].copy(), vals[:nw].copy()), (nodes[:nn].astype(np.int32), start[:nn + 1].copy(),
                                                     feat[:start[nn]].astype(np.int32))


def search_by_bow(fvA, descA, angleA, validA, fvB, descB, angleB, validB, th_low=50, strict=False, nnratio=0.6,
                  check_orientation=True, n_left_b=-1):
    """fvA / fvB = (nodes, start, feat).  -> nmatches, matchA, matchAR."""
    a = [np.ascontiguousarray(x, np.int32) for x in fvA]
    b = [np.ascontiguousarray(x, np.int32) for x in fvB]
    descA = np.ascontiguousarray(descA, np.uint8); descB = np.ascontiguousarray(descB, np.uint8)
    angleA = np.ascontiguousarray(angleA, np.float32); angleB = np.ascontiguousarray(angleB, np.float32)
    validA = np.ascontiguousarray(validA, np.uint8)
    vb = None if validB is None else np.ascontiguousarray(validB, np.uint8)
    nA, nB = len(descA), len(descB)
    mA, mR = np.empty(nA, np.int32), np.empty(nA, np.int32)
    n = lib().oracle_search_by_bow(len(a[0]), _p(a[0]), _p(a[1]), _p(a[2]), _p(descA), _p(angleA), _p(validA), nA,
                                   len(b[0]), _p(b[0]), _p(b[1]), _p(b[2]), _p(descB), _p(angleB),
                                   None if vb is None else _p(vb), nB, int(th_low), int(strict), C.c_float(nnratio),
                                   int(check_orientation), int(n_left_b), _p(mA), _p(mR))
    return n, mA, mR


def search_for_triangulation(fvA, keysA, descA, urightA, has_mp_a, fvB, keysB, descB, urightB, has_mp_b, f12, ep,
                             scale_factors_b, level_sigma2_b, only_stereo=False, coarse=False, check_orientation=True,
                             th_low=50):
    a = [np.ascontiguousarray(x, np.int32) for x in fvA]
    b = [np.ascontiguousarray(x, np.int32) for x in fvB]
    keysA = np.ascontiguousarray(keysA); keysB = np.ascontiguousarray(keysB)
    descA = np.ascontiguousarray(descA, np.uint8); descB = np.ascontiguousarray(descB, np.uint8)
    ura = None if urightA is None else np.ascontiguousarray(urightA, np.float32)
    urb = None if urightB is None else np.ascontiguousarray(urightB, np.float32)
    mpa = np.ascontiguousarray(has_mp_a, np.uint8); mpb = np.ascontiguousarray(has_mp_b, np.uint8)
    f12 = np.ascontiguousarray(f12, np.float32); ep = np.ascontiguousarray(ep, np.float32)
    sf = np.ascontiguousarray(scale_factors_b, np.float32); s2 = np.ascontiguousarray(level_sigma2_b, np.float32)
    m12 = np.empty(len(keysA), np.int32)
    n = lib().oracle_search_for_triangulation(
        len(a[0]), _p(a[0]), _p(a[1]), _p(a[2]), _p(keysA), _p(descA), None if ura is None else _p(ura), _p(mpa), len(keysA),
        len(b[0]), _p(b[0]), _p(b[1]), _p(b[2]), _p(keysB), _p(descB), None if urb is None else _p(urb), _p(mpb), len(keysB),
        _p(f12), _p(ep), _p(sf), _p(s2), int(only_stereo), int(coarse), int(check_orientation), int(th_low), _p(m12))
    return n, m12


def kb8_null_vectors(A):
    """The null-vector step of KannalaBrandt8::Triangulate on n 4x4 float systems -> [n, 4] float64."""
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 16)
    x = np.empty((len(A), 4), np.float64)
    lib().oracle_kb8_null_vectors(_p(A), len(A), _p(x))
    return x


def search_for_triangulation_rig(fvA, keysA, descA, has_mp_a, fvB, keysB, descB, has_mp_b, scale_factors_b, level_sigma2_a,
                                 level_sigma2_b, n_left1, n_left2, pairs, only_stereo=False, coarse=False,
                                 check_orientation=True, th_low=50):
    """Two-camera keyframes (ORBmatcher.cc:1071-1095, 1160-1241).  pairs: float32 [4, 30] = per (bRight1, bRight2)
    combination {P1[8], P2[8], prec1, prec2, R12[9] row-major, t12[3]}."""
    a = [np.ascontiguousarray(x, np.int32) for x in fvA]
    b = [np.ascontiguousarray(x, np.int32) for x in fvB]
    keysA = np.ascontiguousarray(keysA); keysB = np.ascontiguousarray(keysB)
    descA = np.ascontiguousarray(descA, np.uint8); descB = np.ascontiguousarray(descB, np.uint8)
    mpa = np.ascontiguousarray(has_mp_a, np.uint8); mpb = np.ascontiguousarray(has_mp_b, np.uint8)
    sf = np.ascontiguousarray(scale_factors_b, np.float32)
    s2a = np.ascontiguousarray(level_sigma2_a, np.float32); s2b = np.ascontiguousarray(level_sigma2_b, np.float32)
    pairs = np.ascontiguousarray(pairs, np.float32).reshape(4, 30)
    m12 = np.empty(len(keysA), np.int32)
    n = lib().oracle_search_for_triangulation_rig(
        len(a[0]), _p(a[0]), _p(a[1]), _p(a[2]), _p(keysA), _p(descA), _p(mpa), len(keysA),
        len(b[0]), _p(b[0]), _p(b[1]), _p(b[2]), _p(keysB), _p(descB), _p(mpb), len(keysB),
        _p(sf), _p(s2a), _p(s2b), int(n_left1), int(n_left2), _p(pairs), int(only_stereo), int(coarse), int(check_orientation),
        int(th_low), _p(m12))
    return n, m12


# ---------------------------------------------------------------- frame intake
def cvt_gray(img, rgb=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w, c = img.shape
    out = np.empty((h, w), np.uint8)
    lib().oracle_cvt_gray(_p(img), w, h, img.strides[0], c, int(rgb), _p(out), w)
    return out


def remap_linear(src, mapx, mapy):
    src = np.ascontiguousarray(src, np.uint8)
    mapx = np.ascontiguousarray(mapx, np.float32); mapy = np.ascontiguousarray(mapy, np.float32)
    dh, dw = mapx.shape
    out = np.empty((dh, dw), np.uint8)
    lib().oracle_remap_linear(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(mapx), _p(mapy), dw, dh, _p(out), dw)
    return out


def undistort_points(xy, K, dist):
    """cv::undistortPoints(xy, K, dist, None, K) for n x 2 float points; K = (fx, fy, cx, cy) as float32 values."""
    xy = np.ascontiguousarray(xy, np.float32)
    d = np.ascontiguousarray(np.asarray(dist, np.float32).astype(np.float64))
    out = np.empty_like(xy)
    fx, fy, cx, cy = [float(np.float32(v)) for v in K]
    lib().oracle_undistort_points(_p(xy), len(xy), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy), _p(d),
                                  len(d), _p(out))
    return out


def distinctive_descriptors(desc, start):
    desc = np.ascontiguousarray(desc, np.uint8)
    start = np.ascontiguousarray(start, np.int32)
    best = np.empty(len(start) - 1, np.int32)
    lib().oracle_distinctive_descriptors(_p(desc), _p(start), len(start) - 1, _p(best))
    return best


# ---- KannalaBrandt8 geometry behind the fisheye stereo matcher (oracle/kb8_oracle.cpp) ----
def kb8_project(params, p3d):
    p3d = np.ascontiguousarray(p3d, np.float32)
    uv = np.empty((len(p3d), 2), np.float32)
    lib().oracle_kb8_project(_p(np.ascontiguousarray(params, np.float32)), _p(p3d), len(p3d), _p(uv))
    return uv


def kb8_unproject(params, uv, precision=1e-6):
    uv = np.ascontiguousarray(uv, np.float32)
    rays = np.empty((len(uv), 3), np.float32)
    lib().oracle_kb8_unproject(_p(np.ascontiguousarray(params, np.float32)), C.c_float(precision), _p(uv), len(uv), _p(rays))
    return rays


def kb8_triangulate(params1, params2, R12, t12, pt1, pt2, sigma1, unc2, precision1=1e-6, precision2=1e-6):
    """KannalaBrandt8::TriangulateMatches per match: (depth or the negative rejection code, p3D (NaN where rejected))."""
    pt1, pt2 = np.ascontiguousarray(pt1, np.float32), np.ascontiguousarray(pt2, np.float32)
    n = len(pt1)
    depth = np.empty(n, np.float32)
    p3d = np.full((n, 3), np.nan, np.float32)
    lib().oracle_kb8_triangulate(_p(np.ascontiguousarray(params1, np.float32)), C.c_float(precision1),
                                 _p(np.ascontiguousarray(params2, np.float32)), C.c_float(precision2),
                                 _p(np.ascontiguousarray(R12, np.float32)), _p(np.ascontiguousarray(t12, np.float32)),
                                 _p(pt1), _p(pt2), _p(np.ascontiguousarray(sigma1, np.float32)),
                                 _p(np.ascontiguousarray(unc2, np.float32)), n, _p(depth), _p(p3d))
    return depth, p3d
