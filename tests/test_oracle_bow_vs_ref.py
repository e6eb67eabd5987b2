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
