"""Pins oracle/bow_oracle.cpp to DBoW2 itself (Thirdparty/DBoW2 of the reference compiled verbatim: whole
BowVector.cpp, FeatureVector.cpp, FORB.cpp, ScoringObject.cpp, TemplatedVocabulary.h) and to the reference's own
ORBmatcher::SearchByBoW bodies (oracle/_ref/libref_orbmatcher.so, see oracle/ref_build.sh).  CPU only."""
import numpy as np
import pytest

import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.matcher_available(), reason="oracle/_ref/libref_orbmatcher.so not built")


@pytest.fixture(scope="module")
def voc_pair(tmp_path_factory):
    voc = synth.make_vocabulary(10, 4, 3)
    path = str(tmp_path_factory.mktemp("voc") / "voc.txt")
    synth.write_vocabulary_text(path, voc)
    return voc, R.RefVocabulary(path), O.Vocabulary(10, 4, voc["parent"], voc["desc"], voc["weight"])


def test_vocabulary_loads_like_dbow2(voc_pair):
    voc, rv, ov = voc_pair
    assert rv.size() == 10 ** 4


@pytest.mark.parametrize("levelsup", [2, 1, 4, 7])
def test_transform_vs_dbow2(voc_pair, levelsup):
    voc, rv, ov = voc_pair
    desc = np.concatenate([synth.descriptors_near_words(voc, 1500, 5), synth.random_descriptors(500, 6)])
    rw, rwt, rn = rv.transform_features(desc, levelsup)
    ow, owt, on = ov.transform_features(desc, levelsup)
    assert np.array_equal(rw, ow) and np.array_equal(rwt.view(np.uint64), owt.view(np.uint64)) and np.array_equal(rn, on)
    (rid, rval), (rnode, rstart, rfeat) = rv.transform(desc, levelsup)
    (oid, oval), (onode, ostart, ofeat) = ov.transform(desc, levelsup)
    assert np.array_equal(rid, oid) and np.array_equal(rval.view(np.uint64), oval.view(np.uint64))
    assert np.array_equal(rnode, onode) and np.array_equal(rstart, ostart) and np.array_equal(rfeat, ofeat)
    assert len(rid) > 500 and abs(rval.sum() - 1.0) < 1e-9 and (rwt == 0).sum() > 0


@pytest.mark.parametrize("scoring,weighting", [(1, 0), (5, 1), (0, 2), (5, 3)])
def test_transform_other_weightings(tmp_path, scoring, weighting):
    voc = synth.make_vocabulary(6, 3, 8, early_leaf_frac=0.2)
    path = str(tmp_path / "v.txt")
    synth.write_vocabulary_text(path, voc, scoring, weighting)
    rv = R.RefVocabulary(path)
    ov = O.Vocabulary(6, 3, voc["parent"], voc["desc"], voc["weight"], scoring, weighting)
    desc = synth.descriptors_near_words(voc, 800, 2, noise=30)
    for levelsup in (1, 2):       # early leaves sit on level L-1: with levelsup >= 1 the node level is at or above them
                                  # (below it the reference leaves *nid uninitialised)
        (rid, rval), rf = rv.transform(desc, levelsup)
        (oid, oval), of = ov.transform(desc, levelsup)
        assert np.array_equal(rid, oid) and np.array_equal(rval.view(np.uint64), oval.view(np.uint64))
        assert all(np.array_equal(a, b) for a, b in zip(rf, of))


def _bow_frames(voc, n_kf, n_f, seed, n_left=-1):
    """A keyframe and a frame observing the same scene: frame descriptors = noisy copies of keyframe ones."""
    from oracle.oracle import KP_DTYPE
    rng = np.random.default_rng(seed)
    dk = synth.descriptors_near_words(voc, n_kf, seed + 1)
    kk = np.zeros(n_kf, KP_DTYPE)
    kk["x"], kk["y"] = rng.uniform(20, 730, n_kf), rng.uniform(20, 460, n_kf)
    kk["angle"] = rng.uniform(0, 360, n_kf)
    src = rng.integers(0, n_kf, n_f)
    df = np.stack([synth.flip_bits(dk[s], int(rng.integers(0, 25)), rng) for s in src])
    df[::11] = synth.random_descriptors(len(df[::11]), seed + 2)
    kf_ = np.zeros(n_f, KP_DTYPE)
    kf_["x"], kf_["y"] = rng.uniform(20, 730, n_f), rng.uniform(20, 460, n_f)
    kf_["angle"] = (kk["angle"][src] + 25 + rng.normal(0, 10, n_f)) % 360.0
    return kk, dk, kf_, df, rng


SF = np.cumprod(np.concatenate([[1.0], np.full(7, 1.2)])).astype(np.float32)


@pytest.mark.parametrize("seed,check_ori,nnratio", [(1, True, 0.7), (2, False, 0.9), (3, True, 0.6)])
def test_search_by_bow_kf_frame_vs_reference(voc_pair, seed, check_ori, nnratio):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), ORBmatcher.cc:260-494 (Nleft == -1)."""
    voc, rv, ov = voc_pair
    kk, dk, kf_, df, rng = _bow_frames(voc, 1500, 1400, seed)
    R.set_bounds((0.0, 0.0, 752.0, 480.0))
    KF, F = R.RefFrame(kk, dk, SF), R.RefFrame(kf_, df, SF)
    state = rng.choice(3, len(kk), p=[0.2, 0.7, 0.1])            # keyframe slot: no point / good point / bad point
    KF.set_mappoints(state > 0, bad=(state == 2))
    R.compute_bow(KF, rv, 2); R.compute_bow(F, rv, 2)
    rn, rout = R.search_by_bow_kf_f(KF, F, nnratio, check_ori)
    _, fva = ov.transform(dk, 2)
    _, fvb = ov.transform(df, 2)
    n, mA, _ = O.search_by_bow(fva, dk, kk["angle"], state == 1, fvb, df, kf_["angle"], None, 50, False, nnratio, check_ori)
    out = np.full(len(kf_), -1, np.int32)
    out[mA[mA >= 0]] = np.flatnonzero(mA >= 0)                   # vpMapPointMatches[bestIdxF] = pMP of KF slot iA
    assert rn == n and n > 150
    assert np.array_equal(rout, out)


@pytest.mark.parametrize("seed,check_ori", [(4, True), (5, False)])
def test_search_by_bow_kf_frame_fisheye_vs_reference(voc_pair, seed, check_ori):
    """The F.Nleft != -1 branch (:322-343, :374-407): left and right bests per keyframe point."""
    from oracle.oracle import KP_DTYPE
    voc, rv, ov = voc_pair
    kk, dk, kf_, df, rng = _bow_frames(voc, 1200, 1600, seed)
    nl = 900
    l2r = np.full(nl, -1, np.int32); r2l = np.full(len(kf_) - nl, -1, np.int32)
    R.set_bounds((0.0, 0.0, 752.0, 480.0))
    KF = R.RefFrame(kk, dk, SF)
    F = R.RefFrame(kf_[:nl], df[:nl], SF, right=(kf_[nl:], df[nl:], l2r, r2l))
    state = rng.choice(3, len(kk), p=[0.2, 0.7, 0.1])
    KF.set_mappoints(state > 0, bad=(state == 2))
    R.compute_bow(KF, rv, 2); R.compute_bow(F, rv, 2)
    rn, rout = R.search_by_bow_kf_f(KF, F, 0.7, check_ori)
    _, fva = ov.transform(dk, 2)
    _, fvb = ov.transform(df, 2)
    n, mA, mR = O.search_by_bow(fva, dk, kk["angle"], state == 1, fvb, df, kf_["angle"], None, 50, False, 0.7, check_ori, nl)
    out = np.full(len(kf_), -1, np.int32)
    out[mA[mA >= 0]] = np.flatnonzero(mA >= 0)
    out[mR[mR >= 0]] = np.flatnonzero(mR >= 0)
    assert rn == n and (mR >= 0).sum() > 30 and (mA >= 0).sum() > 30
    assert np.array_equal(rout, out)


@pytest.mark.parametrize("seed,check_ori", [(6, True), (7, False)])
def test_search_by_bow_kf_kf_vs_reference(voc_pair, seed, check_ori):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), ORBmatcher.cc:893-1044."""
    voc, rv, ov = voc_pair
    k1, d1, k2, d2, rng = _bow_frames(voc, 1500, 1400, seed)
    R.set_bounds((0.0, 0.0, 752.0, 480.0))
    KF1, KF2 = R.RefFrame(k1, d1, SF), R.RefFrame(k2, d2, SF)
    s1 = rng.choice(3, len(k1), p=[0.2, 0.7, 0.1]); s2 = rng.choice(3, len(k2), p=[0.2, 0.7, 0.1])
    KF1.set_mappoints(s1 > 0, bad=(s1 == 2)); KF2.set_mappoints(s2 > 0, bad=(s2 == 2))
    R.compute_bow(KF1, rv, 2); R.compute_bow(KF2, rv, 2)
    rn, rout = R.search_by_bow_kf_kf(KF1, KF2, 0.75, check_ori)
    _, fva = ov.transform(d1, 2)
    _, fvb = ov.transform(d2, 2)
    n, mA, _ = O.search_by_bow(fva, d1, k1["angle"], s1 == 1, fvb, d2, k2["angle"], s2 == 1, 50, True, 0.75, check_ori)
    assert rn == n and n > 100
    assert np.array_equal(rout, mA)


def _tri_case(voc, seed, stereo_frac):
    """Two keyframes of one scene related by a sideways translation: KF2's keypoints are KF1's shifted along x by a
    depth-dependent disparity (so the epipolar constraint holds for true matches) plus unrelated ones."""
    from oracle.oracle import KP_DTYPE
    rng = np.random.default_rng(seed)
    n1, n2 = 1500, 1400
    d1 = synth.descriptors_near_words(voc, n1, seed + 1)
    k1 = np.zeros(n1, KP_DTYPE)
    k1["x"], k1["y"] = rng.uniform(20, 730, n1), rng.uniform(20, 460, n1)
    k1["angle"] = rng.uniform(0, 360, n1)
    k1["octave"] = rng.integers(0, 8, n1)
    src = rng.integers(0, n1, n2)
    d2 = np.stack([synth.flip_bits(d1[s], int(rng.integers(0, 25)), rng) for s in src])
    k2 = k1[src].copy()
    k2["x"] = k2["x"] - rng.uniform(2, 60, n2).astype(np.float32)              # along the epipolar line (pure x translation)
    k2["y"] = k2["y"] + rng.normal(0, 0.8, n2).astype(np.float32)
    off = rng.uniform(size=n2) < 0.25                                          # off the epipolar line
    k2["y"][off] += rng.uniform(5, 40, off.sum()).astype(np.float32)
    k2["angle"] = (k2["angle"] + 15 + rng.normal(0, 8, n2)) % 360.0
    k2["octave"] = np.clip(k2["octave"] + rng.integers(-1, 2, n2), 0, 7)
    ur1 = np.where(rng.uniform(size=n1) < stereo_frac, k1["x"] - 10, -1).astype(np.float32)
    ur2 = np.where(rng.uniform(size=n2) < stereo_frac, k2["x"] - 10, -1).astype(np.float32)
    mp1 = rng.uniform(size=n1) < 0.3
    mp2 = rng.uniform(size=n2) < 0.3
    return k1, d1, ur1, mp1, k2, d2, ur2, mp2


@pytest.mark.parametrize("seed,only_stereo,coarse,check_ori,stereo_frac", [(1, False, False, True, 0.0), (2, False, False, False, 0.5),
                                                                           (3, True, False, True, 0.6), (4, False, True, True, 0.3)])
def test_search_for_triangulation_vs_reference(voc_pair, seed, only_stereo, coarse, check_ori, stereo_frac):
    """ORBmatcher::SearchForTriangulation, ORBmatcher.cc:1046-1324, pinhole keyframes; Pinhole::epipolarConstrain
    (Pinhole.cpp:186-216) compiled from the reference too."""
    voc, rv, ov = voc_pair
    k1, d1, ur1, mp1, k2, d2, ur2, mp2 = _tri_case(voc, seed, stereo_frac)
    R.set_bounds((0.0, 0.0, 752.0, 480.0))
    KF1, KF2 = R.RefFrame(k1, d1, SF, uright=ur1), R.RefFrame(k2, d2, SF, uright=ur2)
    for kf, t in ((KF1, (0.0, 0.0, 0.0)), (KF2, (-0.3, 0.01, 0.02))):
        R.set_camera(kf, 458.654, 457.296, 367.215, 248.375)
        kf.set_pose(t)
    KF1.set_mappoints(mp1); KF2.set_mappoints(mp2)
    R.compute_bow(KF1, rv, 2); R.compute_bow(KF2, rv, 2)
    rn, rm12, f12, ep = R.search_for_triangulation(KF1, KF2, only_stereo, coarse, check_ori)
    assert rn >= 0
    _, fva = ov.transform(d1, 2)
    _, fvb = ov.transform(d2, 2)
    n, m12 = O.search_for_triangulation(fva, k1, d1, ur1, mp1, fvb, k2, d2, ur2, mp2, f12, ep, SF, SF * SF, only_stereo, coarse,
                                        check_ori)
    assert rn == n and n > 60
    assert np.array_equal(rm12, m12)
    if not coarse:   # the epipolar gate must matter: the coarse search accepts more
        nc, _ = O.search_for_triangulation(fva, k1, d1, ur1, mp1, fvb, k2, d2, ur2, mp2, f12, ep, SF, SF * SF, only_stereo, True,
                                           check_ori)
        assert nc > n
