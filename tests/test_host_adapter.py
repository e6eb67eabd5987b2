"""The header-only C++ adapter (orb-slam3_byzyh_b200/host/) keeps ORB-SLAM3's ORBextractor
signatures: it must compile and link against libORBfe_b200.so (CPU), and produce the oracle's
keypoints / descriptors / bordered pyramid when run (GPU)."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "orb-slam3_byzyh_b200")
EXE = os.path.join(HERE, "native", "host_adapter_check")


def _build():
    so = os.path.join(PKG, "libORBfe_b200.so")
    if not os.path.exists(so):
        pytest.skip("libORBfe_b200.so not built (run __graft_entry__.build())")
    src = os.path.join(HERE, "native", "host_adapter_check.cpp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "oracle", "cvshim"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"), src,
                           os.path.join(ROOT, "oracle", "cvprims.cpp"), "-o", EXE, "-L", PKG, "-lORBfe_b200",
                           "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])
    return EXE


def test_adapter_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_adapter_matches_oracle(tmp_path):
    import synth
    from oracle import oracle as O
    exe = _build()
    img = synth.synth_frame(480, 752, 11)
    raw = tmp_path / "in.raw"
    raw.write_bytes(img.tobytes())
    voc = synth.make_vocabulary(8, 3, 4)
    synth.write_vocabulary_text(str(tmp_path / "voc.txt"), voc)
    out = subprocess.check_output([exe, "480", "752", "1000", "0", "1000", str(raw), str(tmp_path / "o"),
                                   str(tmp_path / "voc.txt")], text=True)
    mono, n, levels, w1, self_d, sf = out.split()
    ex = O.Extractor(1000)
    omono, okps, odesc = ex(img, (0, 1000))
    assert int(mono) == omono and int(n) == len(okps) and int(levels) == 8 and int(self_d) == 0
    assert (tmp_path / "o.kps").read_bytes() == okps.tobytes()
    assert (tmp_path / "o.desc").read_bytes() == odesc.tobytes()
    assert (tmp_path / "o.pyr1").read_bytes() == ex.level(1)["padded"].tobytes()

    # matcher adapter calls made by the same program (see host_adapter_check.cpp)
    n = len(okps)
    sf = ex.tables()["scale"]
    i = np.arange(n)
    pts = dict(u=(okps["x"] + np.float32(1)).astype(np.float32), v=(okps["y"] - np.float32(1)).astype(np.float32),
               ur=(okps["x"] - np.float32(4)).astype(np.float32), radius=(np.float32(6) * sf[okps["octave"]]).astype(np.float32),
               min_level=okps["octave"] - 1, max_level=okps["octave"], angle=okps["angle"],
               valid=((i % 17) != 0).astype(np.uint8), blocks=np.ones(n, np.uint8), desc=odesc)
    uright = np.where(i % 3, okps["x"] - np.float32(5), np.float32(-1)).astype(np.float32)
    bounds = (0.0, 0.0, 752.0, 480.0)
    got = np.frombuffer((tmp_path / "o.match").read_bytes(), np.int32).reshape(4, n)
    inv = ex.tables()["inv_sigma2"]
    fuse, _ = O.search_window(okps, odesc, uright, bounds, pts, 50, True, inv)
    _, sim3 = O.search_by_sim3(okps, odesc, okps, odesc, bounds, pts, pts, 100)
    none = np.zeros(n, np.uint8)
    _, asg, _, _ = O.search_by_projection(okps, odesc, uright, bounds, pts, 0, 100, 0.8, False, none,
                                          np.full(n, -2, np.int32), sf)
    _, asim, _, _ = O.search_by_projection(okps, odesc, None, bounds, pts, 2, 50, 1.0, False, ((i % 5) == 0).astype(np.uint8),
                                           np.full(n, -2, np.int32), sf)
    assert np.array_equal(got[0], fuse) and (fuse >= 0).sum() > 0.5 * n
    assert np.array_equal(got[1], sim3) and (sim3 >= 0).sum() > 0.5 * n
    assert np.array_equal(got[2], asg) and np.array_equal(got[3], asim)

    # vocabulary adapter: loadFromTextFile + transform + SearchByBoW of the frame against itself
    ov = O.Vocabulary(8, 3, voc["parent"], voc["desc"], voc["weight"])
    (wid, wval), fv = ov.transform(odesc, 2)
    rec = np.frombuffer((tmp_path / "o.bow").read_bytes(), np.dtype([("id", "<u4"), ("v", "<f8")]))
    assert np.array_equal(rec["id"], wid) and np.array_equal(rec["v"].view(np.uint64), wval.view(np.uint64))
    _, mA, _ = O.search_by_bow(fv, odesc, okps["angle"], (i % 7) != 0, fv, odesc, okps["angle"], None, 50, False, 0.9, True)
    got = np.frombuffer((tmp_path / "o.bowmatch").read_bytes(), np.int32)
    assert np.array_equal(got, mA) and (mA >= 0).sum() > 0.3 * n
    sf = ex.tables()["scale"]
    has_mp = (i & 1).astype(np.uint8)
    f12 = np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], np.float32)
    nt, m12 = O.search_for_triangulation(fv, okps, odesc, None, has_mp, fv, okps, odesc, None, has_mp, f12,
                                         np.array([-1000, -1000], np.float32), sf, ex.tables()["sigma2"], False, True, True)
    pairs = np.frombuffer((tmp_path / "o.tri").read_bytes(), np.int32).reshape(-1, 2)
    assert np.array_equal(pairs[:, 0], np.flatnonzero(m12 >= 0)) and np.array_equal(pairs[:, 1], m12[m12 >= 0]) and nt > 100
