"""GPU parity of the bag-of-words path (csrc/bow.cu) against oracle/bow_oracle.cpp, which
tests/test_oracle_bow_vs_ref.py pins to DBoW2 itself and to the reference's SearchByBoW."""
import os
import sys

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


@pytest.fixture(scope="module")
def vocs(orbfe):
    voc = synth.make_vocabulary(10, 4, 3)
    return voc, orbfe.ORBVocabulary(10, 4, voc["parent"], voc["desc"], voc["weight"]), \
        O.Vocabulary(10, 4, voc["parent"], voc["desc"], voc["weight"])


@pytest.mark.parametrize("levelsup", [2, 0, 4, 9])
def test_transform_small_vocabulary(vocs, levelsup):
    voc, gv, ov = vocs
    desc = np.concatenate([synth.descriptors_near_words(voc, 3000, 5), synth.random_descriptors(1000, 6)])
    gw, gwt, gn = gv.transform_features(desc, levelsup)
    ow, owt, on = ov.transform_features(desc, levelsup)
    assert np.array_equal(gw, ow) and np.array_equal(gwt.view(np.uint64), owt.view(np.uint64)) and np.array_equal(gn, on)
    (gid, gval), gfv = gv.transform(desc, levelsup)
    (oid, oval), ofv = ov.transform(desc, levelsup)
    assert np.array_equal(gid, oid) and np.array_equal(gval.view(np.uint64), oval.view(np.uint64))
    assert all(np.array_equal(a, b) for a, b in zip(gfv, ofv))


def test_transform_orbvoc_sized_vocabulary(orbfe):
    """k = 10, L = 6 (the shape of ORBvoc.txt: 1 111 111 nodes, 10^6 words), levelsup = 4 as in Frame::ComputeBoW."""
    voc = synth.make_vocabulary_fast(10, 6, 1)
    gv = orbfe.ORBVocabulary(10, 6, voc["parent"], voc["desc"], voc["weight"])
    ov = O.Vocabulary(10, 6, voc["parent"], voc["desc"], voc["weight"])
    rng = np.random.default_rng(2)
    leaves = rng.integers(len(voc["parent"]) - 10 ** 6, len(voc["parent"]), 4000)
    desc = np.stack([synth.flip_bits(voc["desc"][p], int(rng.integers(0, 20)), rng) for p in leaves])
    gw, gwt, gn = gv.transform_features(desc, 4)
    ow, owt, on = ov.transform_features(desc, 4)
    assert np.array_equal(gw, ow) and np.array_equal(gwt.view(np.uint64), owt.view(np.uint64)) and np.array_equal(gn, on)
    assert len(np.unique(gn)) > 50 and gw.max() < 10 ** 6


@pytest.mark.parametrize("weighting,scoring", [(1, 1), (2, 0), (3, 5)])
def test_transform_unbalanced_other_weightings(orbfe, weighting, scoring):
    voc = synth.make_vocabulary(6, 3, 8, early_leaf_frac=0.2)
    gv = orbfe.ORBVocabulary(6, 3, voc["parent"], voc["desc"], voc["weight"], scoring, weighting)
    ov = O.Vocabulary(6, 3, voc["parent"], voc["desc"], voc["weight"], scoring, weighting)
    desc = synth.descriptors_near_words(voc, 900, 2, noise=30)
    for levelsup in (0, 1, 2):
        gw, gwt, gn = gv.transform_features(desc, levelsup)
        ow, owt, on = ov.transform_features(desc, levelsup)
        assert np.array_equal(gw, ow) and np.array_equal(gn, on)
        (gid, gval), gfv = gv.transform(desc, levelsup)
        (oid, oval), ofv = ov.transform(desc, levelsup)
        assert np.array_equal(gid, oid) and np.array_equal(gval.view(np.uint64), oval.view(np.uint64))
        assert all(np.array_equal(a, b) for a, b in zip(gfv, ofv))


def _frames(voc, n_a, n_b, seed):
    rng = np.random.default_rng(seed)
    da = synth.descriptors_near_words(voc, n_a, seed + 1)
    src = rng.integers(0, n_a, n_b)
    db = np.stack([synth.flip_bits(da[s], int(rng.integers(0, 25)), rng) for s in src])
    db[::11] = synth.random_descriptors(len(db[::11]), seed + 2)
    ang_a = rng.uniform(0, 360, n_a).astype(np.float32)
    ang_b = ((ang_a[src] + 25 + rng.normal(0, 10, n_b)) % 360.0).astype(np.float32)
    return da, ang_a, db, ang_b, rng


@pytest.mark.parametrize("seed,check_ori,nnratio,n_left", [(1, True, 0.7, -1), (2, False, 0.9, -1), (3, True, 0.6, 1000),
                                                           (4, False, 0.75, 700)])
def test_search_by_bow_keyframe_frame(orbfe, vocs, seed, check_ori, nnratio, n_left):
    voc, gv, ov = vocs
    da, ang_a, db, ang_b, rng = _frames(voc, 2000, 1800, seed)
    valid_a = (rng.uniform(size=len(da)) < 0.75).astype(np.uint8)
    _, fva = gv.transform(da, 2)
    _, fvb = gv.transform(db, 2)
    m = orbfe.ORBmatcher(nnratio, check_ori)
    n, mA, mR = m.SearchByBoW((da, ang_a, valid_a, fva), (db, ang_b, None, fvb), n_left)
    en, emA, emR = O.search_by_bow(fva, da, ang_a, valid_a, fvb, db, ang_b, None, 50, False, nnratio, check_ori, n_left)
    assert n == en and n > 150
    assert np.array_equal(mA, emA)
    if n_left != -1:
        assert np.array_equal(mR, emR) and (emR >= 0).sum() > 20


@pytest.mark.parametrize("seed,check_ori", [(6, True), (7, False)])
def test_search_by_bow_keyframes(orbfe, vocs, seed, check_ori):
    voc, gv, ov = vocs
    da, ang_a, db, ang_b, rng = _frames(voc, 2000, 1800, seed)
    valid_a = (rng.uniform(size=len(da)) < 0.75).astype(np.uint8)
    valid_b = (rng.uniform(size=len(db)) < 0.75).astype(np.uint8)
    _, fva = gv.transform(da, 2)
    _, fvb = gv.transform(db, 2)
    m = orbfe.ORBmatcher(0.75, check_ori)
    n, mA = m.SearchByBoWKeyFrames((da, ang_a, valid_a, fva), (db, ang_b, valid_b, fvb))
    en, emA, _ = O.search_by_bow(fva, da, ang_a, valid_a, fvb, db, ang_b, valid_b, 50, True, 0.75, check_ori)
    assert n == en and n > 100 and np.array_equal(mA, emA)
    assert not np.any((mA >= 0) & (valid_a == 0)) and np.all(valid_b[mA[mA >= 0]] == 1)


def test_search_by_bow_empty_and_disjoint(orbfe, vocs):
    voc, gv, ov = vocs
    da, ang_a, db, ang_b, rng = _frames(voc, 300, 280, 9)
    _, fva = gv.transform(da, 2)
    _, fvb = gv.transform(db, 2)
    m = orbfe.ORBmatcher(0.7, True)
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    n, mA, _ = m.SearchByBoW((da, ang_a, None, fva), (db, ang_b, None, empty))
    assert n == 0 and np.all(mA == -1)
    shifted = (fvb[0] + 100000, fvb[1], fvb[2])            # no common node
    n, mA, _ = m.SearchByBoW((da, ang_a, None, fva), (db, ang_b, None, shifted))
    assert n == 0 and np.all(mA == -1)


@pytest.mark.parametrize("seed,only_stereo,coarse,check_ori,stereo_frac", [(1, False, False, True, 0.0), (2, False, False, False, 0.5),
                                                                           (3, True, False, True, 0.6), (4, False, True, True, 0.3)])
def test_search_for_triangulation(orbfe, vocs, seed, only_stereo, coarse, check_ori, stereo_frac):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:1046-1324) + Pinhole::epipolarConstrain, pinhole keyframes."""
    from test_oracle_bow_vs_ref import SF, _tri_case
    voc, gv, ov = vocs
    k1, d1, ur1, mp1, k2, d2, ur2, mp2 = _tri_case(voc, seed, stereo_frac)
    # F12 for K = (458.654, 457.296, 367.215, 248.375), R12 = I, t12 = (0.3, -0.01, -0.02); epipole of camera 1 in image 2
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
    t = np.array([0.3, -0.01, -0.02])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    f12 = (np.linalg.inv(K).T @ tx @ np.linalg.inv(K)).astype(np.float32)
    c2 = -t
    ep = np.array([fx * c2[0] / c2[2] + cx, fy * c2[1] / c2[2] + cy], np.float32)
    _, fva = gv.transform(d1, 2)
    _, fvb = gv.transform(d2, 2)
    m = orbfe.ORBmatcher(0.6, check_ori)
    n, m12 = m.SearchForTriangulation((k1, d1, ur1, mp1, fva), (k2, d2, ur2, mp2, fvb), f12, ep, SF, SF * SF, only_stereo, coarse)
    en, em12 = O.search_for_triangulation(fva, k1, d1, ur1, mp1, fvb, k2, d2, ur2, mp2, f12, ep, SF, SF * SF, only_stereo, coarse,
                                          check_ori)
    assert n == en and n > 60 and np.array_equal(m12, em12)


def _rig_keyframes(voc, seed, npts=1400):
    """Two KannalaBrandt8 stereo-rig keyframes observing one cloud of 3-D points: every point is seen by some of the four
    cameras (left / right of keyframe 1 and 2) at its projection (+ noise, a quarter far off), with descriptors that are
    noisy copies of one word-near descriptor per point.  Returns the two sides in the reference's layout
    ([mvKeys | mvKeysRight]) and the four (bRight1, bRight2) camera pairs with their relative poses."""
    from oracle.oracle import KP_DTYPE
    from test_oracle_kb8 import P1, P2
    rng = np.random.default_rng(seed)

    def rot(ax, ang):
        c, s = np.cos(ang), np.sin(ang)
        R = np.eye(3)
        i, j = [(1, 2), (0, 2), (0, 1)][ax]
        R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
        return R
    Rlr, tlr = rot(1, 0.02), np.array([0.1, 0.002, -0.001])        # right -> left of a rig: X_l = Rlr X_r + tlr
    Rk, tk = rot(1, -0.05) @ rot(0, 0.01), np.array([0.35, -0.02, 0.05])   # keyframe 2 left -> keyframe 1 left
    X1l = np.stack([rng.uniform(-2, 2, npts), rng.uniform(-1.5, 1.5, npts), rng.uniform(0.6, 4, npts)], 1)
    cams = {"1l": X1l, "1r": (X1l - tlr) @ Rlr}                  # Rlr^T (X - tlr)
    X2l = (X1l - tk) @ Rk
    cams["2l"], cams["2r"] = X2l, (X2l - tlr) @ Rlr
    base = synth.descriptors_near_words(voc, npts, seed + 1)
    sides = {}
    for kf in ("1", "2"):
        keys, desc, src = [], [], []
        nleft = 0
        for cam, P in (("l", P1), ("r", P2)):
            seen = np.flatnonzero(rng.uniform(size=npts) < 0.6)
            uv = O.kb8_project(P, np.ascontiguousarray(cams[kf + cam][seen], np.float32))
            uv = uv + rng.normal(0, 0.4, uv.shape).astype(np.float32)
            off = rng.uniform(size=len(seen)) < 0.25
            uv[off, 1] += rng.uniform(6, 40, off.sum()).astype(np.float32)
            k = np.zeros(len(seen), KP_DTYPE)
            k["x"], k["y"] = uv[:, 0], uv[:, 1]
            k["angle"] = rng.uniform(0, 360, len(seen))
            k["octave"] = rng.integers(0, 8, len(seen))
            keys.append(k)
            desc.append(np.stack([synth.flip_bits(base[s], int(rng.integers(0, 20)), rng) for s in seen]))
            src.append(seen)
            if cam == "l":
                nleft = len(seen)
        sides[kf] = (np.concatenate(keys), np.concatenate(desc), nleft, np.concatenate(src))
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    pairs = [(P1, P1, 1e-6, 1e-6, f32(Rk), f32(tk)),                                                  # ll
             (P1, P2, 1e-6, 1e-6, f32(Rk @ Rlr), f32(Rk @ tlr + tk)),                                 # lr
             (P2, P1, 1e-6, 1e-6, f32(Rlr.T @ Rk), f32(Rlr.T @ (tk - tlr))),                          # rl
             (P2, P2, 1e-6, 1e-6, f32(Rlr.T @ Rk @ Rlr), f32(Rlr.T @ (Rk @ tlr + tk - tlr)))]         # rr
    return sides, pairs


@pytest.mark.parametrize("seed,coarse,check_ori", [(1, False, False), (2, False, True), (3, True, True)])
def test_search_for_triangulation_two_camera_keyframes(orbfe, vocs, seed, coarse, check_ori):
    """The mpCamera2 branch of ORBmatcher::SearchForTriangulation (ORBmatcher.cc:1071-1095, 1160-1241): per (bRight1,
    bRight2) combination the cameras and the relative pose change and the gate is KannalaBrandt8::epipolarConstrain.
    Against the oracle's restatement; KB8's float transcendentals make the gate tolerance-based (tests/test_gpu_kb8.py), so
    a handful of matches may differ -- the bar is > 99.5 % identical entries; with bCoarse (no gate) everything is exact."""
    from test_oracle_bow_vs_ref import SF
    voc, gv, ov = vocs
    sides, pairs = _rig_keyframes(voc, seed)
    (k1, d1, nl1, s1), (k2, d2, nl2, s2) = sides["1"], sides["2"]
    rng = np.random.default_rng(seed + 9)
    mp1, mp2 = rng.uniform(size=len(k1)) < 0.2, rng.uniform(size=len(k2)) < 0.2
    _, fva = gv.transform(d1, 2)
    _, fvb = gv.transform(d2, 2)
    S2 = (SF * SF).astype(np.float32)
    flat = np.stack([np.concatenate([p[0], p[1], [p[2], p[3]], p[4].reshape(-1), p[5]]) for p in pairs]).astype(np.float32)
    en, em12 = O.search_for_triangulation_rig(fva, k1, d1, mp1, fvb, k2, d2, mp2, SF, S2, S2, nl1, nl2, flat, False, coarse, check_ori)
    m = orbfe.ORBmatcher(0.6, check_ori)
    n, m12 = m.SearchForTriangulation((k1, d1, None, mp1, fva), (k2, d2, None, mp2, fvb), np.zeros(9, np.float32), np.zeros(2, np.float32),
                                      SF, S2, False, coarse, rig=dict(n_left1=nl1, n_left2=nl2, level_sigma2_1=S2, pairs=pairs))
    assert en > 150
    if coarse:
        assert n == en and np.array_equal(m12, em12)
    else:
        assert abs(n - en) <= 3 and (m12 == em12).mean() > 0.995
        good = (m12 >= 0) & (m12 == em12)
        assert (s1[good] == s2[m12[good]]).mean() > 0.9          # the gate keeps true correspondences
    # all four camera combinations occur among the matches
    r1, r2 = np.flatnonzero(m12 >= 0) >= nl1, m12[m12 >= 0] >= nl2
    assert len({(bool(a), bool(b)) for a, b in zip(r1, r2)}) == 4


def test_bow_fold_on_the_device(orbfe, vocs):
    """Frame::ComputeBoW for a batch of frames without leaving the device: descent + fold (orbfe_bow_fold_device) against the
    oracle's transform(features, v, fv, levelsup) (pinned to DBoW2 compiled verbatim): BowVector word ids and values (double
    bit patterns: the additions run in DBoW2's order), FeatureVector nodes / features; stopped words, repeated words, an
    empty frame, a frame of one feature and one at the 4096-feature capacity."""
    import torch
    voc, gv, ov = vocs
    rng = np.random.default_rng(3)
    sizes = [1500, 0, 1, 700, 4096, 2300]
    frames = []
    for k, n in enumerate(sizes):
        d = np.concatenate([synth.descriptors_near_words(voc, n - n // 4, 50 + k), synth.random_descriptors(n // 4, 60 + k)]) if n else np.zeros((0, 32), np.uint8)
        if n > 10:
            d[5] = d[2]; d[9] = d[2]                       # the same word several times: weights add up in feature order
        frames.append(np.ascontiguousarray(d[rng.permutation(len(d))] if n else d))
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    d_desc = torch.from_numpy(np.concatenate(frames)).cuda()
    for levelsup in (2, 4):
        out = {k: v.cpu().numpy() for k, v in gv.transform_batch_device(d_desc, start, 4096, levelsup).items()}
        for b, d in enumerate(frames):
            (ids, vals), (nodes, fstart, feat) = ov.transform(d, levelsup)
            nb, nf = out["n_bow"][b], out["n_fv"][b]
            assert nb == len(ids) and nf == len(nodes), (b, levelsup)
            assert np.array_equal(out["bow_word"][b, :nb].view(np.uint32), ids)
            assert np.array_equal(out["bow_value"][b, :nb].view(np.uint64), vals.view(np.uint64)), (b, levelsup)
            assert np.array_equal(out["fv_node"][b, :nf], nodes) and np.array_equal(out["fv_start"][b, :nf + 1], fstart)
            assert np.array_equal(out["fv_feat"][b, :fstart[-1]], feat)
