"""Argument validation, empty inputs and concurrent use of the matcher-side C ABI (the reference calls ORBmatcher from the
tracking, local-mapping and loop-closing threads at the same time)."""
import ctypes as C
import os
import sys
import threading

import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def orbfe():
    sys.path.insert(0, os.path.join(ROOT, "orb-slam3_byzyh_b200"))
    import orbfe as m
    return m


def test_invalid_arguments_are_reported_not_crashed(orbfe):
    L = orbfe.lib()
    from orbfe._lib import ERR_INVALID, BowSide, FrameView, ProjPoints, TriParams, TriSide, WindowParams
    fv, pp, wp = FrameView(), ProjPoints(), WindowParams()
    out = np.zeros(4, np.int32)
    assert L.orbfe_search_window(None, None, None, None, None, 0) == ERR_INVALID
    fv.n, pp.m = 5, 3                      # sizes without arrays
    assert L.orbfe_search_window(C.byref(fv), C.byref(pp), C.byref(wp), out.ctypes.data, None, 0) == ERR_INVALID
    wp.gate = 7
    assert L.orbfe_search_window(C.byref(fv), C.byref(pp), C.byref(wp), out.ctypes.data, None, 0) == ERR_INVALID
    assert L.orbfe_search_window(C.byref(fv), C.byref(pp), C.byref(wp), out.ctypes.data, None, 99) == ERR_INVALID
    assert L.orbfe_search_by_sim3(C.byref(fv), C.byref(fv), C.byref(pp), C.byref(pp), 100, out.ctypes.data, 0) == ERR_INVALID
    a, b = BowSide(), BowSide()
    a.n = b.n = 4
    a.fv.n_nodes = b.fv.n_nodes = 2
    assert L.orbfe_search_by_bow(C.byref(a), C.byref(b), 50, 0, 0.7, 1, -1, out.ctypes.data, None, 0) == ERR_INVALID
    assert L.orbfe_search_by_bow(C.byref(a), C.byref(b), 50, 0, 0.7, 1, 2, out.ctypes.data, None, 0) == ERR_INVALID
    t, prm = TriSide(), TriParams()
    t.n, t.fv.n_nodes, prm.n_levels = 4, 2, 8
    assert L.orbfe_search_for_triangulation(C.byref(t), C.byref(t), C.byref(prm), out.ctypes.data, 0) == ERR_INVALID
    h = C.c_void_p()
    assert L.orbfe_vocabulary_create(10, 6, 1, None, None, None, 0, C.byref(h)) == ERR_INVALID and not h.value
    par = np.array([0, 5, 0], np.int32)    # parent id out of range
    d, w = np.zeros((3, 32), np.uint8), np.ones(3)
    assert L.orbfe_vocabulary_create(2, 1, 3, par.ctypes.data, d.ctypes.data, w.ctypes.data, 0, C.byref(h)) == ERR_INVALID
    img = np.zeros((4, 4, 2), np.uint8)
    assert L.orbfe_cvt_gray(img.ctypes.data, 4, 4, 8, 2, 0, out.ctypes.data, 4, 0) == ERR_INVALID          # 2 channels
    assert L.orbfe_remap_linear(img.ctypes.data, 4, 4, 2, None, None, 4, 4, out.ctypes.data, 4, 0) == ERR_INVALID
    assert L.orbfe_resize_linear(img.ctypes.data, 4, 4, 4, 0, 4, out.ctypes.data, 4, 0) == ERR_INVALID
    start = np.array([0, 3, 2], np.int32)  # descending range
    assert L.orbfe_distinctive_descriptors(d.ctypes.data, start.ctypes.data, 2, out.ctypes.data, 0) == ERR_INVALID
    assert "ascending" in orbfe.last_error()


def test_empty_inputs(orbfe):
    m = orbfe.ORBmatcher()
    d = synth.map_vs_frame(50, 40, 1)
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"])
    empty = {k: np.zeros(0, t) for k, t in [("u", np.float32), ("v", np.float32), ("radius", np.float32), ("min_level", np.int32),
                                            ("max_level", np.int32), ("valid", np.uint8)]}
    empty["desc"] = np.zeros((0, 32), np.uint8)
    n, bi, bd = m.FuseSearch(F, empty)
    assert n == 0 and len(bi) == 0
    E = orbfe.FrameData(d["keys"][:0], d["fdesc"][:0], d["bounds"])
    pts = dict(u=d["u"], v=d["v"], radius=np.full(50, 10, np.float32), min_level=np.zeros(50, np.int32),
               max_level=np.full(50, 7, np.int32), valid=np.ones(50, np.uint8), desc=d["mdesc"])
    n, bi, bd = m.FuseSearch(E, pts)
    assert n == 0 and np.all(bi == -1) and np.all(bd == 256)
    assert len(orbfe.ORBmatcher.ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), np.zeros(1, np.int32))) == 0
    assert list(orbfe.ORBmatcher.ComputeDistinctiveDescriptors(d["fdesc"][:1], np.array([0, 0, 1], np.int32))) == [-1, 0]
    voc = synth.make_vocabulary(4, 2, 0)
    gv = orbfe.ORBVocabulary(4, 2, voc["parent"], voc["desc"], voc["weight"])
    (ids, vals), (nodes, start, feat) = gv.transform(np.zeros((0, 32), np.uint8), 1)
    assert len(ids) == 0 and len(nodes) == 0 and list(start) == [0]


def test_concurrent_matcher_calls_from_threads(orbfe):
    """Three host threads issue different matchers at once (per-thread staging arenas and streams): every result equals
    the serial one."""
    d = synth.map_vs_frame(4000, 1500, 3)
    rng = np.random.default_rng(0)
    sf = d["scale_factors"]
    lvl = d["level"]
    pts = dict(u=d["u"], v=d["v"], ur=d["u"], radius=(np.float32(4) * sf[lvl]).astype(np.float32),
               min_level=(lvl - 1).astype(np.int32), max_level=lvl.astype(np.int32), angle=np.zeros(4000, np.float32),
               valid=np.ones(4000, np.uint8), blocks=np.ones(4000, np.uint8), desc=d["mdesc"])
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"])
    voc = synth.make_vocabulary(10, 3, 1)
    gv = orbfe.ORBVocabulary(10, 3, voc["parent"], voc["desc"], voc["weight"])
    q = synth.random_descriptors(1500, 4)
    t = synth.random_descriptors(20000, 5)
    cl, asg = np.zeros(1500, np.uint8), np.full(1500, -2, np.int32)
    jobs = {
        "proj": lambda: orbfe.ORBmatcher(0.8, True).SearchByProjection(F, pts, cl, asg)[1],
        "fuse": lambda: orbfe.ORBmatcher().FuseSearch(F, pts)[1],
        "knn": lambda: orbfe.ORBmatcher().knn2(q, t)[0],
        "bow": lambda: gv.transform_features(d["fdesc"], 1)[0],
    }
    serial = {k: f() for k, f in jobs.items()}
    errors = []

    def worker(name, fn):
        try:
            for _ in range(15):
                if not np.array_equal(fn(), serial[name]):
                    errors.append(name)
        except Exception as e:      # noqa: BLE001
            errors.append("%s: %r" % (name, e))
    threads = [threading.Thread(target=worker, args=kv) for kv in jobs.items()]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
