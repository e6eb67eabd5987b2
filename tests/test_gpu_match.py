"""GPU parity of the Hamming matchers (C ABI) against the CPU oracle: DescriptorDistance, kNN-2 +
ratio, the three SearchByProjection modes (incl. sequential claim semantics and the rotation
histogram), rectified stereo matching with SAD sub-pixel refinement.  All outputs bit-exact."""
import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orbfe():
    import orbfe as m
    m.lib()
    return m


def test_descriptor_distance(orbfe):
    a, b = synth.random_descriptors(5000, 1), synth.random_descriptors(5000, 2)
    b[:10] = a[:10]
    b[10] = ~a[10]
    got = orbfe.ORBmatcher.DescriptorDistance(a, b)
    exp = np.unpackbits(a ^ b, axis=1).sum(1)
    assert np.array_equal(got, exp) and got[0] == 0 and got[10] == 256
    assert orbfe.ORBmatcher.DescriptorDistance(a[3], b[77]) == O.hamming(a[3], b[77])


# sizes on both sides of the switch from the scalar kernel to the tensor-core kernel (csrc/knn_umma.cu: nt >= 512 and
# nq * nt >= 2^18), ragged against its 128-query / 256-point tiles, and one map large enough for several tiles per CTA
@pytest.mark.parametrize("nq,nt", [(1, 1), (3, 2), (300, 1), (1500, 1500), (257, 4097), (2000, 20000), (40, 0), (511, 513), (129, 2049),
                                   (777, 100001), (2000, 250000)])
def test_knn2_vs_oracle(orbfe, nq, nt):
    rng = np.random.default_rng(nq * 7 + nt)
    q, t = synth.random_descriptors(nq, nq), synth.random_descriptors(nt, nt + 1)
    if nt > 8:
        t[5] = t[3]                 # exact duplicates: tie -> lower train index
        t[nt - 1] = q[0]
        t[2] = q[0]
        for i in range(0, min(nq, 100)):   # near matches so the ratio test fires both ways
            t[int(rng.integers(0, nt))] = synth.flip_bits(q[i], int(rng.integers(0, 80)), rng)
    m = orbfe.ORBmatcher()
    idx, dist, match = m.knn2(q, t)
    em, eidx, edist = O.fisheye_matches(q, t)
    assert np.array_equal(idx, eidx) and np.array_equal(dist, edist) and np.array_equal(match, em)


def test_knn2_ragged_sizes_on_the_tensor_core_kernel(orbfe):
    """Sixteen random problem sizes that are ragged against the 128-query / 256-point tiles of csrc/knn_umma.cu (every
    one above its size switch), with duplicates of query rows planted at both ends of the map and ties between map rows:
    idx2 / dist2 / match equal the oracle's (cv2 BFMatcher order)."""
    rng = np.random.default_rng(2024)
    m = orbfe.ORBmatcher()
    for case in range(16):
        nq, nt = int(rng.integers(33, 700)), int(rng.integers(8200, 30000))
        q, t = synth.random_descriptors(nq, 1000 + case), synth.random_descriptors(nt, 2000 + case)
        t[0] = q[nq - 1]
        t[nt - 1] = q[0]
        t[nt // 2] = t[nt // 2 + 1]                       # equal rows: the lower index first
        for i in range(0, nq, 5):
            t[int(rng.integers(0, nt))] = synth.flip_bits(q[i], int(rng.integers(0, 90)), rng)
        idx, dist, match = m.knn2(q, t)
        em, eidx, edist = O.fisheye_matches(q, t)
        assert np.array_equal(idx, eidx) and np.array_equal(dist, edist) and np.array_equal(match, em), (nq, nt)


def test_knn2_sharded_merge_equals_whole(orbfe):
    q, t = synth.random_descriptors(500, 3), synth.random_descriptors(30000, 4)
    t[100] = q[7]; t[25000] = q[7]
    m = orbfe.ORBmatcher()
    whole = m.knn2(q, t)
    G = 4
    bounds = np.linspace(0, len(t), G + 1).astype(int)
    parts = [m.knn2(q, t[bounds[g]:bounds[g + 1]], train_offset=int(bounds[g])) for g in range(G)]
    idx, dist, match = m.knn2_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    assert np.array_equal(idx, whole[0]) and np.array_equal(dist, whole[1]) and np.array_equal(match, whole[2])


def _pts_from_map(d, n_map, th, sf, rng, with_angle=True):
    lvl = d["level"][:n_map]
    r = np.where(d["view_cos"][:n_map] > 0.998, 2.5, 4.0).astype(np.float32)
    if th != 1.0:
        r = (r * np.float32(th)).astype(np.float32)
    radius = (r * sf[lvl]).astype(np.float32)
    angle = rng.uniform(0, 360, n_map)
    src = d["src"][:n_map]
    has = src >= 0      # true matches rotate coherently (+35 deg), so the histogram keeps most of them
    angle[has] = (d["keys"]["angle"][src[has]] + 35.0 + rng.normal(0, 12, has.sum())) % 360.0
    return dict(u=d["u"][:n_map], v=d["v"][:n_map], ur=(d["u"][:n_map] - 5).astype(np.float32), radius=radius,
                min_level=(lvl - 1).astype(np.int32), max_level=lvl.astype(np.int32),
                angle=angle.astype(np.float32),
                valid=(rng.uniform(size=n_map) < 0.97).astype(np.uint8),
                blocks=(rng.uniform(size=n_map) < 0.8).astype(np.uint8), desc=d["mdesc"][:n_map])


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n_map,n_frame,seed", [(3000, 800, 1), (60000, 2000, 2)])
def test_search_by_projection(orbfe, mode, n_map, n_frame, seed):
    d = synth.map_vs_frame(n_map, n_frame, seed)
    rng = np.random.default_rng(seed + 50)
    sf = d["scale_factors"]
    pts = _pts_from_map(d, n_map, 3.0, sf, rng)
    if mode == 1:   # last-frame search uses forward/backward octave windows
        fw = rng.uniform(size=n_map) < 0.3
        pts["min_level"] = np.where(fw, d["level"][:n_map], d["level"][:n_map] - 1).astype(np.int32)
        pts["max_level"] = np.where(fw, -1, d["level"][:n_map] + 1).astype(np.int32)
    uright = np.where(rng.uniform(size=n_frame) < 0.5, d["keys"]["x"] - 5 + rng.normal(0, 3, n_frame), -1).astype(np.float32)
    claimed = (rng.uniform(size=n_frame) < 0.05).astype(np.uint8)
    assigned = np.full(n_frame, -1, np.int32)
    assigned[claimed > 0] = 10 ** 6
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], uright)
    m = orbfe.ORBmatcher(nnratio=0.8, checkOri=True)
    th_acc = 100 if mode < 2 else 90
    fn = [m.SearchByProjection, m.SearchByProjectionLastFrame, lambda *a: m.SearchByProjectionKeyFrame(*a, 90)][mode]
    n, asg, bi, bd = fn(F, pts, claimed, assigned)
    en, easg, ebi, ebd = O.search_by_projection(d["keys"], d["fdesc"], uright, d["bounds"], pts, mode, th_acc, 0.8,
                                                True, claimed, assigned, sf)
    assert n == en and n > 0.2 * min(n_frame, n_map)
    assert np.array_equal(bi, ebi) and np.array_equal(bd, ebd) and np.array_equal(asg, easg)


def test_search_conflict_chains(orbfe):
    """Many map points compete for few keypoints: deep claim chains must resolve like the loop."""
    rng = np.random.default_rng(5)
    n_frame, n_map = 40, 4000
    d = synth.map_vs_frame(n_map, n_frame, 9, w=200, h=150)
    base = d["fdesc"][rng.integers(0, n_frame, n_map)]
    mdesc = np.stack([synth.flip_bits(b, int(rng.integers(0, 50)), rng) for b in base])
    pts = dict(u=rng.uniform(0, 200, n_map).astype(np.float32), v=rng.uniform(0, 150, n_map).astype(np.float32),
               ur=np.zeros(n_map, np.float32), radius=np.full(n_map, 60, np.float32),
               min_level=np.zeros(n_map, np.int32), max_level=np.full(n_map, -1, np.int32),
               angle=np.zeros(n_map, np.float32), valid=np.ones(n_map, np.uint8),
               blocks=(rng.uniform(size=n_map) < 0.5).astype(np.uint8), desc=mdesc)
    claimed = np.zeros(n_frame, np.uint8)
    assigned = np.full(n_frame, -1, np.int32)
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], None)
    m = orbfe.ORBmatcher(nnratio=0.9, checkOri=False)
    for mode, fn in enumerate([m.SearchByProjection, m.SearchByProjectionLastFrame]):
        n, asg, bi, bd = fn(F, pts, claimed, assigned)
        en, easg, ebi, ebd = O.search_by_projection(d["keys"], d["fdesc"], None, d["bounds"], pts, mode, 100, 0.9,
                                                    False, claimed, assigned, d["scale_factors"])
        assert n == en and np.array_equal(bi, ebi) and np.array_equal(bd, ebd) and np.array_equal(asg, easg)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_stereo_matches(orbfe, seed):
    """C2: EuRoC-shaped rectified pair, nFeatures 1200, lapping {0,0}."""
    left, right = synth.stereo_pair(480, 752, seed)
    gl, gr = orbfe.ORBextractor(1200), orbfe.ORBextractor(1200)
    cl, cr = O.Extractor(1200), O.Extractor(1200)
    _, kl, dl = gl(left, None, (0, 0))
    _, kr, dr = gr(right, None, (0, 0))
    _, okl, odl = cl(left, (0, 0))
    _, okr, odr = cr(right, (0, 0))
    assert kl.tobytes() == okl.tobytes() and kr.tobytes() == okr.tobytes()
    mbf, fx = 47.90639384423901, 435.2046959714599
    mb = mbf / fx
    ur, dp = orbfe.ORBmatcher.ComputeStereoMatches(gl, gr, kl, dl, kr, dr, mbf, mb)
    eur, edp = O.stereo_match(cl, cr, okl, odl, okr, odr, mbf, mb)
    assert (eur > 0).sum() > 200
    assert np.array_equal(ur.view(np.uint32), eur.view(np.uint32))
    assert np.array_equal(dp.view(np.uint32), edp.view(np.uint32))


def test_fisheye_stereo_c3(orbfe):
    """C3: 512x512 pair, nFeatures 1500, lapping area {0,511}: kNN-2 + 0.7 ratio on the stereo halves."""
    a, b = synth.shifted_pair(512, 512, 3)
    ga, gb = orbfe.ORBextractor(1500), orbfe.ORBextractor(1500)
    ma, ka, da = ga(a, None, (0, 511))
    mb_, kb, db = gb(b, None, (0, 511))
    oa = O.Extractor(1500)(a, (0, 511))
    ob = O.Extractor(1500)(b, (0, 511))
    assert ma == oa[0] and mb_ == ob[0] and np.array_equal(da, oa[2]) and np.array_equal(db, ob[2])
    m = orbfe.ORBmatcher()
    idx, dist, match = m.knn2(da[ma:], db[mb_:])       # Frame.cc:1534-1553
    em, eidx, edist = O.fisheye_matches(oa[2][oa[0]:], ob[2][ob[0]:])
    assert np.array_equal(match, em) and np.array_equal(idx, eidx) and np.array_equal(dist, edist)
    assert (match >= 0).sum() > 100


def test_c5_full_size_search_and_knn(orbfe):
    """C5 at BASELINE size: 1 M map points vs 2000 frame descriptors.  SearchByProjection against the
    oracle (fast on CPU: few candidates per point); brute-force kNN-2 through size-independent
    properties (sharded == whole) plus a numpy spot check of 24 queries."""
    n_map, n_frame = 1000000, 2000
    d = synth.map_vs_frame(n_map, n_frame, 5)
    rng = np.random.default_rng(77)
    pts = _pts_from_map(d, n_map, 1.0, d["scale_factors"], rng)
    claimed = np.zeros(n_frame, np.uint8)
    assigned = np.full(n_frame, -1, np.int32)
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], None)
    m = orbfe.ORBmatcher(nnratio=0.8, checkOri=False)
    n, asg, bi, bd = m.SearchByProjection(F, pts, claimed, assigned)
    en, easg, ebi, ebd = O.search_by_projection(d["keys"], d["fdesc"], None, d["bounds"], pts, 0, 100, 0.8, False,
                                                claimed, assigned, d["scale_factors"])
    assert n == en and n > 1000
    assert np.array_equal(bi, ebi) and np.array_equal(bd, ebd) and np.array_equal(asg, easg)
    # brute force: whole map vs 8 shards merged (the multi-GPU decomposition), then spot check
    idx, dist, match = m.knn2(d["fdesc"], d["mdesc"])
    b = np.linspace(0, n_map, 9).astype(int)
    parts = [m.knn2(d["fdesc"], d["mdesc"][b[g]:b[g + 1]], train_offset=int(b[g])) for g in range(8)]
    midx, mdist, mmatch = m.knn2_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    assert np.array_equal(midx, idx) and np.array_equal(mdist, dist) and np.array_equal(mmatch, match)
    lut = np.array([bin(i).count("1") for i in range(256)], np.int32)
    for q in rng.choice(n_frame, 24, replace=False):
        dd = lut[d["mdesc"] ^ d["fdesc"][q]].sum(1)
        order = np.lexsort((np.arange(n_map), dd))[:2]
        assert list(idx[q]) == list(order) and list(dist[q]) == [int(dd[order[0]]), int(dd[order[1]])]
    has = np.flatnonzero(d["src"] >= 0)
    true_of = {int(d["src"][j]): j for j in has[::-1]}
    hit = sum(1 for q in range(n_frame) if q in true_of and dist[q, 0] <= 60)
    assert hit > 0.95 * n_frame      # every frame descriptor has a planted near copy in the map


def _fisheye_case(n_map, nl, nr, seed, all_block):
    """A fisheye-stereo-shaped frame: left / right keypoint sets with their own descriptors, a partial
    mvLeftToRightMatch table, map points projected into both cameras."""
    rng = np.random.default_rng(seed)
    d = synth.map_vs_frame(n_map, nl, seed, w=512, h=512)
    keysL, descL = d["keys"], d["fdesc"]
    # right camera: a shifted, noisy copy of part of the left set plus unrelated keypoints
    keysR = np.zeros(nr, orbfe_kp())
    descR = synth.random_descriptors(nr, seed + 9)
    src = rng.choice(nl, size=nr // 2, replace=False)
    keysR[:nr // 2] = keysL[src]
    keysR["x"][:nr // 2] -= 25.0
    for i, s_ in enumerate(src):
        descR[i] = synth.flip_bits(descL[s_], int(rng.integers(0, 30)), rng)
    keysR["x"][nr // 2:] = rng.uniform(10, 500, nr - nr // 2)
    keysR["y"][nr // 2:] = rng.uniform(10, 500, nr - nr // 2)
    keysR["octave"][nr // 2:] = rng.integers(0, 8, nr - nr // 2)
    l2r = np.full(nl, -1, np.int32)
    r2l = np.full(nr, -1, np.int32)
    for i, s_ in enumerate(src):
        if rng.uniform() < 0.6:
            l2r[s_] = i
            r2l[i] = s_
    sf = d["scale_factors"]
    ptsL = _pts_from_map(d, n_map, 3.0, sf, rng)
    if all_block:
        ptsL["blocks"] = np.ones(n_map, np.uint8)
    ptsR = dict(u=(ptsL["u"] - 25.0 + rng.normal(0, 1, n_map)).astype(np.float32), v=ptsL["v"].copy(),
                radius=ptsL["radius"].copy(), min_level=ptsL["min_level"].copy(), max_level=ptsL["max_level"].copy(),
                valid=(rng.uniform(size=n_map) < 0.8).astype(np.uint8))
    return d, keysL, descL, keysR, descR, l2r, r2l, ptsL, ptsR


def orbfe_kp():
    from oracle.oracle import KP_DTYPE
    return KP_DTYPE


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("all_block", [True, False])
def test_search_by_projection_fisheye(orbfe, mode, all_block):
    """Nleft != -1 branches: left + right camera searches, stereo partner writes, shared rotation
    histogram; with non-blocking points and partner links the ordered single-thread kernel is used."""
    n_map, nl, nr = 6000, 900, 800
    d, keysL, descL, keysR, descR, l2r, r2l, ptsL, ptsR = _fisheye_case(n_map, nl, nr, 21 + mode, all_block)
    rng = np.random.default_rng(3)
    claimed = (rng.uniform(size=nl + nr) < 0.05).astype(np.uint8)
    assigned = np.full(nl + nr, -1, np.int32)
    assigned[claimed > 0] = 10 ** 6
    m = orbfe.ORBmatcher(nnratio=0.8, checkOri=True)
    FL = orbfe.FrameData(keysL, descL, d["bounds"], None)
    FR = orbfe.FrameData(keysR, descR, d["bounds"], None)
    n, asg, bl, br = m.SearchByProjectionFisheye(FL, FR, l2r, r2l, ptsL, ptsR, claimed, assigned, mode)
    en, easg, ebl, ebr = O.search_by_projection_fisheye(keysL, descL, keysR, descR, d["bounds"], l2r, r2l, ptsL, ptsR,
                                                        mode, 100, 0.8, True, claimed, assigned)
    assert n == en and n > 300 and (ebr >= 0).sum() > 100
    assert np.array_equal(bl, ebl) and np.array_equal(br, ebr) and np.array_equal(asg, easg)


def test_sharded_map_device_path(orbfe):
    """orbfe.dist.ShardedMap (device-resident shard, packed gather layout) at world size 1 equals knn2."""
    import torch
    import orbfe.dist as D
    q, t = synth.random_descriptors(700, 5), synth.random_descriptors(50000, 6)
    t[123] = q[9]; t[40000] = q[9]
    dev = torch.device("cuda:0")
    smap = D.ShardedMap(t, 0, dev)
    idx, dist, match = smap.knn2(torch.from_numpy(q).to(dev))
    eidx, edist, ematch = orbfe.ORBmatcher().knn2(q, t)
    assert np.array_equal(idx.cpu().numpy(), eidx) and np.array_equal(dist.cpu().numpy(), edist)
    assert np.array_equal(match.cpu().numpy(), ematch)


@pytest.mark.parametrize("seed", [0, 1])
def test_search_for_initialization(orbfe, seed):
    """Monocular initialisation matcher (ORBmatcher.cc:735-891) on two real extractions of a shifted scene:
    take-over of already matched keypoints, ratio test, rotation histogram, vbPrevMatched update."""
    a, b = synth.shifted_pair(480, 752, seed)
    ex = orbfe.ORBextractor(5000)                      # mpIniORBextractor = 5 x nFeatures
    _, k1, d1 = ex(a, None, (0, 1000))
    _, k2, d2 = ex(b, None, (0, 1000))
    bounds = (0.0, 0.0, 752.0, 480.0)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    m = orbfe.ORBmatcher(nnratio=0.9, checkOri=True)
    n, m12, pv = m.SearchForInitialization(orbfe.FrameData(k1, d1, bounds), orbfe.FrameData(k2, d2, bounds), prev, 100)
    en, em12, epv = O.search_for_initialization(k1, d1, k2, d2, bounds, prev, 100, 0.9, True)
    assert n == en and n > 100
    assert np.array_equal(m12, em12) and np.array_equal(pv.view(np.uint32), epv.view(np.uint32))


def _kf_points(d, n_map, rng, th):
    """Keyframe-side searches: level window [nPredictedLevel-1, nPredictedLevel], radius = th * scale."""
    sf = d["scale_factors"]
    lvl = d["level"][:n_map]
    return dict(u=d["u"][:n_map], v=d["v"][:n_map], ur=(d["u"][:n_map] - 5 + rng.normal(0, 1.5, n_map)).astype(np.float32),
                radius=(np.float32(th) * sf[lvl]).astype(np.float32), min_level=(lvl - 1).astype(np.int32),
                max_level=lvl.astype(np.int32), valid=(rng.uniform(size=n_map) < 0.9).astype(np.uint8),
                desc=d["mdesc"][:n_map])


@pytest.mark.parametrize("gate", [True, False])
@pytest.mark.parametrize("n_map,n_frame,seed", [(2500, 1200, 3), (40000, 3000, 4), (7, 1, 5)])
def test_fuse_search(orbfe, gate, n_map, n_frame, seed):
    """Inner loop of both ORBmatcher::Fuse overloads (ORBmatcher.cc:1326-1534 with the chi2 gate, :1536-1688 without)."""
    d = synth.map_vs_frame(n_map, n_frame, seed)
    rng = np.random.default_rng(seed + 11)
    pts = _kf_points(d, n_map, rng, 3.0)
    uright = np.where(rng.uniform(size=n_frame) < 0.5, d["keys"]["x"] - 5 + rng.normal(0, 1.5, n_frame), -1).astype(np.float32)
    sf = d["scale_factors"]
    inv_sigma2 = (np.float32(1.0) / (sf * sf)).astype(np.float32) if gate else None
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], uright)
    n, bi, bd = orbfe.ORBmatcher().FuseSearch(F, pts, inv_sigma2)
    ebi, ebd = O.search_window(d["keys"], d["fdesc"], uright, d["bounds"], pts, 50, gate, inv_sigma2)
    assert np.array_equal(bi, ebi) and np.array_equal(bd, ebd) and n == (ebi >= 0).sum()
    if n_map > 100:
        assert n > 0.1 * min(n_map, n_frame)
        if gate:   # the gate must actually reject candidates the ungated search keeps
            ubi, _ = O.search_window(d["keys"], d["fdesc"], uright, d["bounds"], pts, 50, False, None)
            assert (ubi != ebi).sum() > 0


@pytest.mark.parametrize("seed", [0, 1])
def test_search_by_sim3(orbfe, seed):
    """ORBmatcher.cc:1690-1940: both directions + mutual agreement."""
    from oracle.oracle import KP_DTYPE
    rng = np.random.default_rng(seed)
    n1, n2 = 1800, 2100
    d = synth.map_vs_frame(n1, n2, seed + 30)          # KF2 keypoints = d["keys"]; KF1's map points = the "map"
    sf = d["scale_factors"]
    k2, d2 = d["keys"], d["fdesc"]
    k1 = np.zeros(n1, KP_DTYPE)
    k1["x"], k1["y"], k1["octave"] = d["u"], d["v"], d["level"]
    d1 = d["mdesc"]
    pts12 = _kf_points(d, n1, rng, 7.5)                # KF1 point i1 projected into KF2
    # reverse direction: KF2's point i2 projects near the KF1 keypoint it came from (if any), else anywhere
    back = np.full(n2, -1)
    src = d["src"]
    for i1 in range(n1):
        if src[i1] >= 0:
            back[src[i1]] = i1
    u21 = rng.uniform(0, 1280, n2).astype(np.float32); v21 = rng.uniform(0, 720, n2).astype(np.float32)
    has = back >= 0
    u21[has] = k1["x"][back[has]] + rng.normal(0, 2, has.sum()); v21[has] = k1["y"][back[has]] + rng.normal(0, 2, has.sum())
    lvl2 = np.clip(k2["octave"] + (rng.uniform(size=n2) < 0.2), 0, 7).astype(np.int32)
    pts21 = dict(u=u21, v=v21, radius=(np.float32(7.5) * sf[lvl2]).astype(np.float32), min_level=lvl2 - 1, max_level=lvl2,
                 valid=(rng.uniform(size=n2) < 0.9).astype(np.uint8), desc=d2)
    F1, F2 = orbfe.FrameData(k1, d1, d["bounds"]), orbfe.FrameData(k2, d2, d["bounds"])
    n, m12 = orbfe.ORBmatcher().SearchBySim3(F1, F2, pts12, pts21)
    en, em12 = O.search_by_sim3(k1, d1, k2, d2, d["bounds"], pts12, pts21, 100)
    assert n == en and n > 200 and np.array_equal(m12, em12)


@pytest.mark.parametrize("ratio", [1.0, 1.5, 0.77])
def test_search_by_projection_sim3(orbfe, ratio):
    """ORBmatcher.cc:496-610: keyframe-side search that skips keypoints already in vpMatched and accepts
    bestDist <= TH_LOW*ratioHamming."""
    n_map, n_frame = 5000, 1500
    d = synth.map_vs_frame(n_map, n_frame, 8)
    rng = np.random.default_rng(80)
    pts = _kf_points(d, n_map, rng, 8.0)
    matched = (rng.uniform(size=n_frame) < 0.2).astype(np.uint8)
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"])
    n, asg, bi, bd = orbfe.ORBmatcher().SearchByProjectionSim3(F, pts, matched, np.full(n_frame, -2, np.int32), ratio)
    th = int(np.floor(np.float32(50) * np.float32(ratio)))
    full = dict(pts, angle=np.zeros(n_map, np.float32), blocks=np.ones(n_map, np.uint8))
    en, easg, ebi, ebd = O.search_by_projection(d["keys"], d["fdesc"], None, d["bounds"], full, 2, th, 1.0, False, matched,
                                                np.full(n_frame, -2, np.int32), d["scale_factors"])
    assert n == en and n > 300
    assert np.array_equal(asg, easg) and np.array_equal(bi, ebi) and np.array_equal(bd, ebd)
    assert not np.any((asg >= 0) & (matched > 0))


@pytest.mark.parametrize("seed", [0, 1])
def test_distinctive_descriptors(orbfe, seed):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:438-529), batched: best row per map point."""
    from test_oracle_match_vs_ref import _distinctive_case, flatten_good_observations
    desc, kf_start, rows, kf_bad = _distinctive_case(seed, n_points=2000)
    good, start = flatten_good_observations(desc, kf_start, rows, kf_bad)
    best = orbfe.ORBmatcher.ComputeDistinctiveDescriptors(good, start)
    assert np.array_equal(best, O.distinctive_descriptors(good, start))
    assert (best == -1).sum() >= 1 and best.max() > 20


# ---- map-sharded SearchByProjection (BASELINE config 5, second half) -------------------------------------------------
def _sharded_search_in_process(orbfe, F, pts, bounds_of_shards, claimed, assigned, nnratio):
    """Several orbfe.dist.ShardedProjection shards of one map driven from ONE process: every pass runs all shards
    against the same claim table (their atomicMin updates into one table ARE the elementwise minimum)."""
    import torch
    import orbfe.dist as D
    dev = torch.device("cuda:0")
    shards = []
    for lo, hi in bounds_of_shards:
        sp = D.ShardedProjection({k: (v[lo:hi] if v is not None else None) for k, v in pts.items()}, lo, dev, nnratio=nnratio)
        sp.set_frame(F)
        shards.append(sp)
    n = len(F.keys)
    static = torch.where(torch.from_numpy(claimed).to(dev) != 0, -1, D.INT_MAX).to(torch.int32)

    def run_pass(cin, cout):
        for sp in shards:
            sp.run_pass(cin, cout)
    claims, passes = D.claim_fixpoint(run_pass, static, lambda c: c)
    mine = torch.full((n,), -1, dtype=torch.int32, device=dev)
    nm = torch.zeros(1, dtype=torch.int32, device=dev)
    for sp in shards:
        sp.finish(mine, nm)
    out = torch.where(mine >= 0, mine, torch.from_numpy(assigned).to(dev))
    bi = np.concatenate([sp.results()[0] for sp in shards])
    bd = np.concatenate([sp.results()[1] for sp in shards])
    return int(nm.item()), out.cpu().numpy(), bi, bd, passes


@pytest.mark.parametrize("n_map,n_frame,seed,cuts", [(3000, 800, 1, (0.5,)), (60000, 2000, 2, (0.1, 0.55, 0.56)),
                                                     (200000, 2000, 3, (0.25, 0.5, 0.75))])
def test_sharded_search_by_projection_equals_single_call(orbfe, n_map, n_frame, seed, cuts):
    """The map split into contiguous shards (unequal, one nearly empty) gives the assignments, per-point results and
    match count of orbfe_search_by_projection on the whole map -- and of the CPU oracle."""
    d = synth.map_vs_frame(n_map, n_frame, seed)
    rng = np.random.default_rng(seed + 50)
    pts = _pts_from_map(d, n_map, 3.0, d["scale_factors"], rng)
    uright = np.where(rng.uniform(size=n_frame) < 0.5, d["keys"]["x"] - 5 + rng.normal(0, 3, n_frame), -1).astype(np.float32)
    claimed = (rng.uniform(size=n_frame) < 0.05).astype(np.uint8)
    assigned = np.full(n_frame, -1, np.int32)
    assigned[claimed > 0] = 10 ** 6
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], uright)
    edges = [0] + [int(c * n_map) for c in cuts] + [n_map]
    n, asg, bi, bd, passes = _sharded_search_in_process(orbfe, F, pts, list(zip(edges[:-1], edges[1:])), claimed, assigned, 0.8)
    m = orbfe.ORBmatcher(nnratio=0.8, checkOri=True)
    en, easg, ebi, ebd = m.SearchByProjection(F, pts, claimed, assigned)
    assert n == en and n > 0.2 * min(n_frame, n_map) and passes >= 2
    assert np.array_equal(asg, easg) and np.array_equal(bi, ebi) and np.array_equal(bd, ebd)
    if n_map <= 60000:
        on, oasg, obi, obd = O.search_by_projection(d["keys"], d["fdesc"], uright, d["bounds"], pts, 0, 100, 0.8, True,
                                                    claimed, assigned, d["scale_factors"])
        assert n == on and np.array_equal(asg, oasg) and np.array_equal(bi, obi) and np.array_equal(bd, obd)


def test_sharded_search_conflict_chains(orbfe):
    """Deep claim chains that cross shard boundaries (4000 points fight for 40 keypoints, 5 shards)."""
    rng = np.random.default_rng(5)
    n_frame, n_map = 40, 4000
    d = synth.map_vs_frame(n_map, n_frame, 9, w=200, h=150)
    base = d["fdesc"][rng.integers(0, n_frame, n_map)]
    mdesc = np.stack([synth.flip_bits(b, int(rng.integers(0, 50)), rng) for b in base])
    pts = dict(u=rng.uniform(0, 200, n_map).astype(np.float32), v=rng.uniform(0, 150, n_map).astype(np.float32),
               ur=np.zeros(n_map, np.float32), radius=np.full(n_map, 60, np.float32),
               min_level=np.zeros(n_map, np.int32), max_level=np.full(n_map, -1, np.int32),
               angle=np.zeros(n_map, np.float32), valid=np.ones(n_map, np.uint8),
               blocks=(rng.uniform(size=n_map) < 0.5).astype(np.uint8), desc=mdesc)
    claimed = np.zeros(n_frame, np.uint8)
    assigned = np.full(n_frame, -1, np.int32)
    F = orbfe.FrameData(d["keys"], d["fdesc"], d["bounds"], None)
    edges = [0, 700, 1500, 1501, 3000, n_map]
    n, asg, bi, bd, passes = _sharded_search_in_process(orbfe, F, pts, list(zip(edges[:-1], edges[1:])), claimed, assigned, 0.9)
    en, easg, ebi, ebd = orbfe.ORBmatcher(nnratio=0.9, checkOri=False).SearchByProjection(F, pts, claimed, assigned)
    assert n == en and passes > 2
    assert np.array_equal(asg, easg) and np.array_equal(bi, ebi) and np.array_equal(bd, ebd)


# ---- batched, device-resident stereo (BASELINE configs 2 and 3 as frame-pair batches) ---------------------------------
def _extract_pairs_device(orbfe, lefts, rights, nf, lap):
    import torch
    from orbfe._lib import KP_DTYPE
    dev = torch.device("cuda:0")
    exL, exR = orbfe.ORBextractor(nf), orbfe.ORBextractor(nf)
    cap = exL.capacity
    st = torch.cuda.Stream(device=dev)
    out = []
    for ex, frames in ((exL, lefts), (exR, rights)):
        B = len(frames)
        d_img = torch.from_numpy(np.stack(frames)).to(dev)
        d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
        d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
        d_n = torch.empty(B, dtype=torch.int32, device=dev)
        d_mono = torch.empty(B, dtype=torch.int32, device=dev)
        ex.extract_batch_device(d_img, lap, d_kps, d_desc, d_n, d_mono, st)
        out.append((d_img, d_kps, d_desc, d_n, d_mono))
    return exL, exR, st, out[0], out[1], KP_DTYPE


def test_stereo_match_batch_device_equals_per_pair_calls(orbfe):
    """64 rectified pairs: orbfe_stereo_match_batch_device on the device-resident pyramids / output slabs equals the
    per-pair host call (orbfe_stereo_match, itself pinned to the reference's ComputeStereoMatches) bit for bit."""
    B = 64
    pairs = [synth.stereo_pair(240, 376, 100 + i) for i in range(B)]
    exL, exR, st, L, R, KP = _extract_pairs_device(orbfe, [p[0] for p in pairs], [p[1] for p in pairs], 600, (0, 0))
    mbf, mb = 47.9, 47.9 / 435.2
    ur, dp = orbfe.ORBmatcher.ComputeStereoMatchesBatchDevice(exL, exR, L[1], L[2], L[3], R[1], R[2], R[3], mbf, mb, st)
    st.synchronize()
    ur, dp = ur.cpu().numpy(), dp.cpu().numpy()
    nL, nR = L[3].cpu().numpy(), R[3].cpu().numpy()
    kL, kR = L[1].cpu().numpy().view(KP).reshape(B, -1), R[1].cpu().numpy().view(KP).reshape(B, -1)
    dL, dR = L[2].cpu().numpy(), R[2].cpu().numpy()
    matched = 0
    for b in range(B):
        eur, edp = orbfe.ORBmatcher.ComputeStereoMatches(exL, exR, kL[b, :nL[b]], dL[b, :nL[b]], kR[b, :nR[b]], dR[b, :nR[b]],
                                                         mbf, mb, frame=b)
        assert np.array_equal(ur[b, :nL[b]].view(np.uint32), eur.view(np.uint32)), b
        assert np.array_equal(dp[b, :nL[b]].view(np.uint32), edp.view(np.uint32)), b
        matched += int((eur >= 0).sum())
    assert matched > 20 * B


@pytest.mark.parametrize("nf", [700, 300])     # >= 512 rows per frame: tensor-core kernel (knn_umma.cu), below: scalar kernel
def test_knn2_batch_device_equals_per_pair_calls(orbfe, nf):
    """32 fisheye-style pairs with a lapping area: the batched kNN-2 + ratio test over rows [mono, n) of both sides equals
    the per-pair orbfe_knn2 (pinned to cv2's BFMatcher) and the CPU oracle."""
    B = 32
    pairs = [synth.shifted_pair(256, 256, 200 + i) for i in range(B)]
    exL, exR, st, L, R, KP = _extract_pairs_device(orbfe, [p[0] for p in pairs], [p[1] for p in pairs], nf, (60, 200))
    idx2, dist2, match = orbfe.ORBmatcher.knn2_batch_device(L[2], L[4], L[3], R[2], R[4], R[3], st)
    st.synchronize()
    idx2, dist2, match = idx2.cpu().numpy(), dist2.cpu().numpy(), match.cpu().numpy()
    nL, nR, mL, mR = (t.cpu().numpy() for t in (L[3], R[3], L[4], R[4]))
    dL, dR = L[2].cpu().numpy(), R[2].cpu().numpy()
    m = orbfe.ORBmatcher()
    total = 0
    for b in range(B):
        q, t = dL[b, mL[b]:nL[b]], dR[b, mR[b]:nR[b]]
        assert len(q) > 50 and len(t) > 50 and mL[b] > 0
        eidx, edist, ematch = m.knn2(q, t)
        k = len(q)
        assert np.array_equal(idx2[b, :k], eidx) and np.array_equal(dist2[b, :k], edist) and np.array_equal(match[b, :k], ematch)
        if b < 4:
            om, oidx, odist = O.fisheye_matches(q, t)
            assert np.array_equal(eidx, oidx) and np.array_equal(edist, odist) and np.array_equal(ematch, om)
        total += int((ematch >= 0).sum())
    assert total > (10 if nf >= 700 else 3) * B
