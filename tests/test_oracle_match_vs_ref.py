"""Pins oracle/match_oracle.cpp (the matcher restatement the CUDA path is checked against) to the
reference's OWN matcher code: ORBmatcher::SearchByProjection (three Frame overloads, incl. the fisheye
branches), SearchForInitialization, DescriptorDistance, Frame::GetFeaturesInArea and
Frame::ComputeStereoMatches, whose function bodies are compiled verbatim from /root/reference into
oracle/_ref/libref_orbmatcher.so (oracle/ref_build.sh).  The conversions from the reference's inputs
(MapPoint members, poses) to the flat arrays of include/orbfe.h follow INTEGRATION.md section 2, so
these tests also check that recipe.  CPU only."""
import numpy as np
import pytest

import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.matcher_available(), reason="oracle/_ref/libref_orbmatcher.so not built")

UNTOUCHED = -2


def _radius_by_viewing_cos(view_cos):
    # ORBmatcher.cc:243-251: float compared with the double literal 0.998
    return np.where(view_cos.astype(np.float64) > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32)


def _same_slots(ref_slots, asg, initial_null):
    """ref: idx / -2 (holds what it held) / -1 (now NULL, was not).  oracle: idx / UNTOUCHED / -1 (culled)."""
    o = asg.copy()
    o[(o == -1) & initial_null] = UNTOUCHED       # culling a slot that was NULL before leaves it as it was
    return np.array_equal(ref_slots, o)


def test_descriptor_distance_vs_reference():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    b[:50] = a[:50]
    b[50:100, :16] = a[50:100, :16]
    a[100], b[100] = 0, 255
    for i in range(300):
        assert R.descriptor_distance(a[i], b[i]) == O.hamming(a[i], b[i])


@pytest.mark.parametrize("seed", [0, 1])
def test_features_in_area_vs_reference(seed):
    d = synth.map_vs_frame(10, 3000, seed)
    bounds = (-13.5, -9.25, 1290.0, 731.5) if seed else d["bounds"]
    R.set_bounds(bounds)
    F = R.RefFrame(d["keys"], d["fdesc"], d["scale_factors"])
    rng = np.random.default_rng(seed + 7)
    for it in range(400):
        x, y = rng.uniform(-80, 1400), rng.uniform(-80, 800)
        r = float(rng.choice([0.5, 3, 10, 40, 150, 2000]))
        lo, hi = [(-1, -1), (0, -1), (2, -1), (0, 3), (3, 3), (-1, 2), (5, 2), (-1, 0)][it % 8]
        ref = F.features_in_area(x, y, r, lo, hi)
        mine = O.features_in_area(d["keys"], bounds, x, y, r, lo, hi)
        assert np.array_equal(ref, mine), (x, y, r, lo, hi)     # same indices in the same (cell-major) order


def _mappoint_case(n_map, n_frame, seed, th, far):
    d = synth.map_vs_frame(n_map, n_frame, seed)
    rng = np.random.default_rng(seed + 50)
    sf = d["scale_factors"]
    lvl = d["level"]
    mp = dict(in_view=(rng.uniform(size=n_map) < 0.95).astype(np.uint8),
              depth=rng.uniform(1, 60, n_map).astype(np.float32),
              bad=(rng.uniform(size=n_map) < 0.03).astype(np.uint8),
              nobs=(rng.uniform(size=n_map) < 0.8).astype(np.int32) * 3,
              proj_x=d["u"], proj_y=d["v"], proj_xr=(d["u"] - 5).astype(np.float32),
              level=lvl, view_cos=d["view_cos"], desc=d["mdesc"])
    mp["view_cos"][::7] = np.float32(0.9995)
    r = _radius_by_viewing_cos(mp["view_cos"])
    if th != 1.0:
        r = (r * np.float32(th)).astype(np.float32)
    valid = (mp["in_view"] > 0) & (mp["bad"] == 0)
    if far:
        valid &= ~(mp["depth"] > np.float32(40.0))
    pts = dict(u=mp["proj_x"], v=mp["proj_y"], ur=mp["proj_xr"], radius=(r * sf[lvl]).astype(np.float32),
               min_level=(lvl - 1).astype(np.int32), max_level=lvl.astype(np.int32), angle=np.zeros(n_map, np.float32),
               valid=valid.astype(np.uint8), blocks=(mp["nobs"] > 0).astype(np.uint8), desc=mp["desc"])
    return d, rng, mp, pts


@pytest.mark.parametrize("n_map,n_frame,seed,th,far", [(3000, 800, 1, 3.0, False), (20000, 2000, 2, 1.0, True),
                                                       (4000, 60, 3, 5.0, False)])
def test_search_mappoints_vs_reference(n_map, n_frame, seed, th, far):
    """ORBmatcher.cc:46-240, rectified-stereo / mono frame (Nleft == -1)."""
    d, rng, mp, pts = _mappoint_case(n_map, n_frame, seed, th, far)
    uright = np.where(rng.uniform(size=n_frame) < 0.5, d["keys"]["x"] - 5 + rng.normal(0, 3, n_frame), -1).astype(np.float32)
    state = rng.choice(3, n_frame, p=[0.9, 0.06, 0.04])      # 0 NULL, 1 map point with observations, 2 without
    R.set_bounds(d["bounds"])
    F = R.RefFrame(d["keys"], d["fdesc"], d["scale_factors"], uright=uright)
    F.set_mappoints(state > 0, nobs=(state == 1).astype(np.int32))
    rn, slots = F.search_mappoints(mp, th, far, 40.0, 0.8)
    claimed = (state == 1).astype(np.uint8)
    n, asg, _, _ = O.search_by_projection(d["keys"], d["fdesc"], uright, d["bounds"], pts, 0, 100, 0.8, False, claimed,
                                          np.full(n_frame, UNTOUCHED, np.int32), d["scale_factors"])
    assert rn == n and n > 0.2 * min(n_map, n_frame)
    assert _same_slots(slots, asg, state == 0)


def _lastframe_case(n_map, n_frame, seed, motion, mono, fisheye=False):
    """Last frame = n_map slots, each optionally holding a map point at world position (u*z, v*z, z) with z a
    power of two, current pose = identity, pinhole fx=fy=1, cx=cy=0: the projection is exactly (u, v)."""
    d = synth.map_vs_frame(n_map, n_frame, seed, w=(512 if fisheye else 1280), h=(512 if fisheye else 720))
    rng = np.random.default_rng(seed + 77)
    sf = d["scale_factors"]
    lvl = d["level"]
    from oracle.oracle import KP_DTYPE
    lkeys = np.zeros(n_map, KP_DTYPE)
    lkeys["x"], lkeys["y"] = d["u"], d["v"]
    lkeys["octave"] = lvl
    lkeys["angle"] = rng.uniform(0, 360, n_map).astype(np.float32)
    src = d["src"]
    has_src = src >= 0     # true matches rotate coherently so the histogram keeps most of them
    lkeys["angle"][has_src] = ((d["keys"]["angle"][src[has_src]] + 35.0 + rng.normal(0, 12, has_src.sum())) % 360.0).astype(np.float32)
    z = rng.choice([0.5, 1.0, 2.0, 4.0, -1.0], n_map, p=[0.2, 0.3, 0.3, 0.17, 0.03]).astype(np.float32)
    u = d["u"].copy()
    u[::41] = np.float32(-3.0)                       # projects outside the image bounds
    xyz = np.stack([u * z, d["v"] * z, z], 1).astype(np.float32)
    has = rng.uniform(size=n_map) < 0.9
    outlier = rng.uniform(size=n_map) < 0.05
    nobs = (rng.uniform(size=n_map) < 0.8).astype(np.int32) * 2
    mb, mbf = np.float32(0.5), np.float32(40.0)
    tz = {"forward": 1.0, "backward": -1.0, "none": 0.1}[motion]
    fwd = motion == "forward" and not mono
    bwd = motion == "backward" and not mono
    W, H = d["bounds"][2], d["bounds"][3]
    invz = (np.float32(1.0) / z).astype(np.float32)
    valid = has & ~outlier & (invz >= 0) & (u >= 0) & (u <= W) & (d["v"] >= 0) & (d["v"] <= H)
    if fwd:
        lo, hi = lvl, np.full(n_map, -1)
    elif bwd:
        lo, hi = np.zeros(n_map), lvl
    else:
        lo, hi = lvl - 1, lvl + 1
    th = 7.0 if mono else 15.0
    pts = dict(u=u, v=d["v"], ur=(u - (mbf * invz).astype(np.float32)).astype(np.float32),
               radius=(np.float32(th) * sf[lvl]).astype(np.float32), min_level=np.asarray(lo, np.int32),
               max_level=np.asarray(hi, np.int32), angle=lkeys["angle"], valid=valid.astype(np.uint8),
               blocks=(nobs > 0).astype(np.uint8), desc=d["mdesc"])
    last = dict(keys=lkeys, has=has, outlier=outlier, nobs=nobs, xyz=xyz, pose=(0.0, 0.0, tz), mb=mb, mbf=mbf, th=th, z=z)
    return d, rng, last, pts


@pytest.mark.parametrize("motion,mono", [("none", False), ("forward", False), ("backward", False), ("forward", True)])
@pytest.mark.parametrize("check_ori", [True, False])
def test_search_lastframe_vs_reference(motion, mono, check_ori):
    """ORBmatcher.cc:1951-2185, Nleft == -1: octave windows by motion, stereo check, rotation histogram."""
    n_map, n_frame = 3000, 900
    d, rng, last, pts = _lastframe_case(n_map, n_frame, 4, motion, mono)
    uright = np.where(rng.uniform(size=n_frame) < 0.5, d["keys"]["x"] - 20 + rng.normal(0, 8, n_frame), -1).astype(np.float32)
    state = rng.choice(3, n_frame, p=[0.9, 0.06, 0.04])
    R.set_bounds(d["bounds"])
    cur = R.RefFrame(d["keys"], d["fdesc"], d["scale_factors"], uright=uright, mb=last["mb"], mbf=last["mbf"])
    cur.set_mappoints(state > 0, nobs=(state == 1).astype(np.int32))
    lf = R.RefFrame(last["keys"], d["mdesc"], d["scale_factors"])
    lf.set_mappoints(last["has"], nobs=last["nobs"], xyz=last["xyz"], desc=d["mdesc"], outlier=last["outlier"])
    lf.set_pose(last["pose"])
    rn, slots = cur.search_lastframe(lf, last["th"], mono, 0.9, check_ori)
    claimed = (state == 1).astype(np.uint8)
    n, asg, _, _ = O.search_by_projection(d["keys"], d["fdesc"], uright, d["bounds"], pts, 1, 100, 0.9, check_ori, claimed,
                                          np.full(n_frame, UNTOUCHED, np.int32), d["scale_factors"])
    assert rn == n and n > 150
    assert _same_slots(slots, asg, state == 0)


@pytest.mark.parametrize("check_ori", [True, False])
def test_search_keyframe_vs_reference(check_ori):
    """ORBmatcher.cc:2197-2325 (relocalisation): distance range, PredictScale window, sAlreadyFound, ORBdist."""
    n_map, n_frame = 4000, 1200
    d, rng, last, pts = _lastframe_case(n_map, n_frame, 6, "none", True)
    sf = d["scale_factors"]
    z = np.abs(last["z"])
    xyz = np.stack([pts["u"] * z, pts["v"] * z, z], 1).astype(np.float32)
    bad = rng.uniform(size=n_map) < 0.03
    found = rng.uniform(size=n_map) < 0.1
    dist3d = np.sqrt((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]) + xyz[:, 2] * xyz[:, 2]).astype(np.float32)
    max_d = (dist3d * rng.uniform(0.7, 4.0, n_map)).astype(np.float32)          # MapPoint::mfMaxDistance
    min_d = (max_d / np.float32(sf[-1]) * rng.uniform(0.5, 1.5, n_map)).astype(np.float32)
    state = rng.choice(3, n_frame, p=[0.9, 0.06, 0.04])
    R.set_bounds(d["bounds"])
    cur = R.RefFrame(d["keys"], d["fdesc"], sf)
    cur.set_mappoints(state > 0, nobs=(state == 1).astype(np.int32))
    kf = R.RefFrame(last["keys"], d["mdesc"], sf)
    kf.set_mappoints(last["has"], nobs=last["nobs"], xyz=xyz, desc=d["mdesc"], bad=bad, min_dist=min_d, max_dist=max_d)
    th, orb_dist = 10.0, 90
    rn, slots = cur.search_keyframe(kf, found, th, orb_dist, 0.9, check_ori)
    # the caller's part: projection, distance range, MapPoint::PredictScale (MapPoint.cc:715-731)
    lvl = np.array([cur.predict_scale(max_d[i], dist3d[i]) for i in range(n_map)], np.int32)
    W, H = d["bounds"][2], d["bounds"][3]
    in_range = ~((dist3d < np.float32(0.8) * min_d) | (dist3d > np.float32(1.2) * max_d))
    valid = last["has"] & ~bad & ~found & (pts["u"] >= 0) & (pts["u"] <= W) & (pts["v"] >= 0) & (pts["v"] <= H) & in_range
    kpts = dict(pts, radius=(np.float32(th) * sf[lvl]).astype(np.float32), min_level=(lvl - 1).astype(np.int32),
                max_level=(lvl + 1).astype(np.int32), valid=valid.astype(np.uint8), blocks=np.ones(n_map, np.uint8))
    claimed = (state > 0).astype(np.uint8)          # :2266 any non-NULL slot is skipped, whatever its observations
    n, asg, _, _ = O.search_by_projection(d["keys"], d["fdesc"], None, d["bounds"], kpts, 2, orb_dist, 0.9, check_ori,
                                          claimed, np.full(n_frame, UNTOUCHED, np.int32), sf)
    assert rn == n and n > 150
    assert _same_slots(slots, asg, state == 0)


def _right_view(d, nl, nr, rng, seed):
    """Right fisheye camera: shifted noisy copies of part of the left set + unrelated keypoints, partial
    mvLeftToRightMatch / mvRightToLeftMatch tables."""
    from oracle.oracle import KP_DTYPE
    keysL, descL = d["keys"], d["fdesc"]
    keysR = np.zeros(nr, KP_DTYPE)
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
    return keysR, descR, l2r, r2l


@pytest.mark.parametrize("all_block", [True, False])
def test_search_mappoints_fisheye_vs_reference(all_block):
    """ORBmatcher.cc:46-240 with Nleft != -1: right-camera branch :171-237, stereo partner writes."""
    n_map, nl, nr = 6000, 900, 800
    d, rng, mp, pts = _mappoint_case(n_map, nl, 21, 3.0, False)
    for k in ("u", "v"):
        d[k] *= np.float32(0.4)            # 1280x720 projections into the 512x512 fisheye image
    d["keys"]["x"] *= np.float32(0.4); d["keys"]["y"] *= np.float32(0.4)
    bounds = (0.0, 0.0, 512.0, 512.0)
    keysR, descR, l2r, r2l = _right_view(d, nl, nr, rng, 21)
    if all_block:
        mp["nobs"][:] = 1
        pts["blocks"] = np.ones(n_map, np.uint8)
    lvl_r = np.where(rng.uniform(size=n_map) < 0.1, -1, np.clip(mp["level"] + rng.integers(-1, 2, n_map), 0, 7)).astype(np.int32)
    mp.update(in_view_r=(rng.uniform(size=n_map) < 0.8).astype(np.uint8), level_r=lvl_r,
              view_cos_r=rng.uniform(0.99, 1.0, n_map).astype(np.float32),
              proj_xr=(mp["proj_x"] - 25.0 + rng.normal(0, 1, n_map)).astype(np.float32), proj_yr=mp["proj_y"].copy())
    mp["in_view"][::9] = 0                  # seen by the right camera only
    ok = (mp["bad"] == 0)
    pts["valid"] = ((mp["in_view"] > 0) & ok).astype(np.uint8)
    sf = d["scale_factors"]
    rr = _radius_by_viewing_cos(mp["view_cos_r"])       # :178 no `th` factor on the right camera
    ptsR = dict(u=mp["proj_xr"], v=mp["proj_yr"], radius=(rr * sf[np.maximum(lvl_r, 0)]).astype(np.float32),
                min_level=(lvl_r - 1).astype(np.int32), max_level=lvl_r.astype(np.int32),
                valid=((mp["in_view_r"] > 0) & ok & (lvl_r != -1)).astype(np.uint8))
    state = rng.choice(3, nl + nr, p=[0.9, 0.06, 0.04])
    R.set_bounds(bounds)
    F = R.RefFrame(d["keys"], d["fdesc"], sf, right=(keysR, descR, l2r, r2l))
    F.set_mappoints(state > 0, nobs=(state == 1).astype(np.int32))
    rn, slots = F.search_mappoints(mp, 3.0, False, 0.0, 0.8)
    claimed = (state == 1).astype(np.uint8)
    n, asg, bl, br = O.search_by_projection_fisheye(d["keys"], d["fdesc"], keysR, descR, bounds, l2r, r2l, pts, ptsR, 0,
                                                    100, 0.8, False, claimed, np.full(nl + nr, UNTOUCHED, np.int32))
    assert rn == n and n > 300 and (br >= 0).sum() > 50
    assert _same_slots(slots, asg, state == 0)


@pytest.mark.parametrize("motion", ["none", "forward", "backward"])
def test_search_lastframe_fisheye_vs_reference(motion):
    """ORBmatcher.cc:1951-2185 with CurrentFrame.Nleft != -1: right-camera projection through Trl (:2085-2155)."""
    n_map, nl, nr = 5000, 900, 800
    d, rng, last, pts = _lastframe_case(n_map, nl, 31, motion, False, fisheye=True)
    bounds = d["bounds"]
    keysR, descR, l2r, r2l = _right_view(d, nl, nr, rng, 31)
    sf = d["scale_factors"]
    state = rng.choice(3, nl + nr, p=[0.9, 0.06, 0.04])
    R.set_bounds(bounds)
    cur = R.RefFrame(d["keys"], d["fdesc"], sf, right=(keysR, descR, l2r, r2l), mb=last["mb"], mbf=last["mbf"])
    cur.set_mappoints(state > 0, nobs=(state == 1).astype(np.int32))
    trl = np.float32(-12.5)
    cur.set_trl((trl, 0.0, 0.0))
    lf = R.RefFrame(last["keys"], d["mdesc"], sf)
    lf.set_mappoints(last["has"], nobs=last["nobs"], xyz=last["xyz"], desc=d["mdesc"], outlier=last["outlier"])
    lf.set_pose(last["pose"])
    rn, slots = cur.search_lastframe(lf, last["th"], False, 0.9, True)
    xyz = last["xyz"]
    with np.errstate(all="ignore"):
        ur = ((xyz[:, 0] + trl).astype(np.float32) / xyz[:, 2]).astype(np.float32)     # project(Trl * x3Dc)
        vr = (xyz[:, 1] / xyz[:, 2]).astype(np.float32)
    lvl = last["keys"]["octave"]
    if motion == "forward":
        lo, hi = lvl, np.full(n_map, -1)
    elif motion == "backward":
        lo, hi = np.zeros(n_map), lvl
    else:
        lo, hi = lvl - 1, lvl + 1
    ptsR = dict(u=ur, v=vr, radius=pts["radius"], min_level=np.asarray(lo, np.int32), max_level=np.asarray(hi, np.int32),
                valid=pts["valid"])
    claimed = (state == 1).astype(np.uint8)
    n, asg, bl, br = O.search_by_projection_fisheye(d["keys"], d["fdesc"], keysR, descR, bounds, l2r, r2l, pts, ptsR, 1,
                                                    100, 0.9, True, claimed, np.full(nl + nr, UNTOUCHED, np.int32))
    assert rn == n and n > 300 and (br >= 0).sum() > 50
    assert _same_slots(slots, asg, state == 0)


@pytest.mark.parametrize("seed,check_ori", [(0, True), (1, False), (2, True)])
def test_search_for_initialization_vs_reference(seed, check_ori):
    """ORBmatcher.cc:735-891: level-0 keypoints only, take-over of already matched keypoints, ratio test,
    rotation histogram, vbPrevMatched update."""
    from oracle.oracle import KP_DTYPE
    rng = np.random.default_rng(seed)
    n1, n2 = 1500, 1700
    k2 = np.zeros(n2, KP_DTYPE)
    k2["x"] = rng.uniform(20, 730, n2); k2["y"] = rng.uniform(20, 460, n2)
    k2["octave"] = rng.choice(4, n2, p=[0.7, 0.15, 0.1, 0.05])
    k2["angle"] = rng.uniform(0, 360, n2)
    d2 = synth.random_descriptors(n2, seed + 3)
    src = rng.integers(0, n2, n1)                  # several F1 keypoints may compete for one F2 keypoint
    k1 = k2[src].copy()
    k1["x"] += rng.normal(8, 10, n1).astype(np.float32)
    k1["y"] += rng.normal(0, 10, n1).astype(np.float32)
    k1["angle"] = (k1["angle"] + 20 + rng.normal(0, 10, n1)) % 360.0
    k1["octave"] = rng.choice(3, n1, p=[0.8, 0.1, 0.1])
    d1 = np.stack([synth.flip_bits(d2[s_], int(rng.integers(0, 70)), rng) for s_ in src])
    bounds = (0.0, 0.0, 752.0, 480.0)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    sf = np.cumprod(np.concatenate([[1.0], np.full(7, 1.2)])).astype(np.float32)
    R.set_bounds(bounds)
    f1, f2 = R.RefFrame(k1, d1, sf), R.RefFrame(k2, d2, sf)
    rn, rm12, rprev = R.search_for_initialization(f1, f2, prev, 100, 0.9, check_ori)
    n, m12, pv = O.search_for_initialization(k1, d1, k2, d2, bounds, prev, 100, 0.9, check_ori)
    assert rn == n and n > 200
    assert np.array_equal(rm12, m12) and np.array_equal(rprev.view(np.uint32), pv.view(np.uint32))


@pytest.mark.parametrize("seed", [0, 1])
def test_stereo_matches_vs_reference(seed):
    """Frame.cc:1102-1358 on a synthetic rectified pair: both images through the reference's own ORBextractor,
    then its ComputeStereoMatches; the restatement gets the same keypoints and its own (pinned) pyramids."""
    L, Rimg = synth.stereo_pair(480, 752, seed)
    mbf, mb = 40.0, 0.11
    kl, dl, kr, dr, ur, dp = R.stereo(L, Rimg, mbf, mb)
    exL, exR = O.Extractor(), O.Extractor()
    _, okl, odl = exL(L)
    _, okr, odr = exR(Rimg)
    assert okl.tobytes() == kl.tobytes() and np.array_equal(odl, dl) and okr.tobytes() == kr.tobytes()
    our, odp = O.stereo_match(exL, exR, kl, dl, kr, dr, mbf, mb)
    assert (ur > 0).sum() > 100
    assert np.array_equal(ur.view(np.uint32), our.view(np.uint32)) and np.array_equal(dp.view(np.uint32), odp.view(np.uint32))


# ---- keyframe-side searches: Fuse x2, SearchByProjection (Sim3), SearchBySim3 --------------------------------
F32 = np.float32
CAM = (F32(458.654), F32(457.296), F32(367.215), F32(248.375))      # EuRoC-like pinhole intrinsics


def _world_points(u, v, lvl, t, rng, sf):
    """World points whose camera-frame position p + t projects near (u, v); MapPoint distance range chosen so
    that PredictScale lands on `lvl`; some points violate the depth / range / normal checks on purpose."""
    m = len(u)
    fx, fy, cx, cy = [float(c) for c in CAM]
    z = rng.uniform(2.0, 12.0, m)
    z[::53] = -1.5
    pc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    xyz = (pc - np.asarray(t, np.float64)).astype(F32)
    po = pc.astype(F32)
    dist = np.linalg.norm(po, axis=1)
    max_d = (dist * 1.2 ** (lvl - 0.5)).astype(F32)
    max_d[::37] *= F32(0.3)                                  # too far for its invariance range
    min_d = (max_d / F32(sf[-1])).astype(F32)
    normal = (po / dist[:, None]).astype(F32)
    flip = rng.uniform(size=m) < 0.05
    normal[flip] *= F32(-1)                                   # seen from behind
    return xyz, normal, min_d, max_d


def _project_kf(xyz, t, with_sim3_scale=None):
    """Float32 mirror of the reference's projection code in Fuse / Fuse(Sim3) / SearchByProjection(Sim3):
    p3Dc = Tcw * p3Dw, uv = project(p3Dc), invz = 1/z, PO = p3Dw - Ow, dist3D = |PO|."""
    fx, fy, cx, cy = CAM
    t = np.asarray(t, F32)
    if with_sim3_scale is not None:
        t = (t / F32(with_sim3_scale)).astype(F32)            # SE3f(Scw.rotationMatrix(), Scw.translation()/Scw.scale())
    pc = (xyz + t).astype(F32)
    with np.errstate(all="ignore"):
        u = (fx * pc[:, 0] / pc[:, 2] + cx).astype(F32)
        v = (fy * pc[:, 1] / pc[:, 2] + cy).astype(F32)
        invz = (F32(1) / pc[:, 2]).astype(F32)
    ow = (-t).astype(F32)
    po = (xyz - ow).astype(F32)
    dist = np.sqrt((po[:, 0] * po[:, 0] + po[:, 1] * po[:, 1]) + po[:, 2] * po[:, 2]).astype(F32)
    return pc, u, v, invz, po, dist


def _kf_valid(pc, u, v, po, dist, P, bounds_int):
    x0, y0, x1, y1 = bounds_int
    dot = ((po[:, 0] * P["normal"][:, 0] + po[:, 1] * P["normal"][:, 1]).astype(F32) + po[:, 2] * P["normal"][:, 2]).astype(F32)
    ok = ~(pc[:, 2] < 0)
    ok &= (u >= x0) & (u < x1) & (v >= y0) & (v < y1)                               # KeyFrame::IsInImage
    ok &= ~((dist < F32(0.8) * P["min_dist"]) | (dist > F32(1.2) * P["max_dist"]))
    ok &= ~(dot.astype(np.float64) < 0.5 * dist.astype(np.float64))
    return ok


def _kf_case(n_map, n_frame, seed, t):
    d = synth.map_vs_frame(n_map, n_frame, seed, w=752, h=480)
    rng = np.random.default_rng(seed + 5)
    sf = d["scale_factors"]
    xyz, normal, min_d, max_d = _world_points(d["u"].astype(np.float64), d["v"].astype(np.float64), d["level"], t, rng, sf)
    P = dict(xyz=xyz, normal=normal, min_dist=min_d, max_dist=max_d, desc=d["mdesc"],
             bad=(rng.uniform(size=n_map) < 0.03).astype(np.uint8))
    return d, rng, sf, P


def _apply_fuse(best_idx, order_ok, slot_mp, slot_bad, slot_obs, cand_obs, sim3):
    """The caller's part of Fuse (ORBmatcher.cc:1480-1501 / :1645-1661) on plain arrays, as in host/ORBmatcher_b200.h."""
    actions, n_fused = [], 0
    repl = np.full(len(best_idx), -1, np.int32)
    slot_mp = list(slot_mp)
    for i in range(len(best_idx)):
        k = best_idx[i]
        if k < 0 or not order_ok[i]:
            continue
        holder = slot_mp[k]
        if holder is not None:
            kind, hid = holder
            h_bad = slot_bad[hid] if kind == "slot" else False
            if not h_bad:
                if sim3:
                    repl[i] = hid if kind == "slot" else -100 - hid
                else:
                    h_id = 1000000 + hid if kind == "slot" else hid
                    h_obs = slot_obs[hid] if kind == "slot" else cand_obs[hid]
                    actions.append((1, i, h_id) if h_obs > cand_obs[i] else (1, h_id, i))
        else:
            actions += [(2, i, k), (3, i, k)]
            slot_mp[k] = ("cand", i)
        n_fused += 1
    return n_fused, np.array(actions, np.int32).reshape(-1, 3), repl


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False)])
def test_fuse_vs_reference(seed, stereo):
    """ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>&, th, bRight=false), ORBmatcher.cc:1326-1534."""
    n_map, n_frame, th = 3000, 1500, 3.0
    t = (0.21, -0.13, 0.4)
    d, rng, sf, P = _kf_case(n_map, n_frame, seed, t)
    P.update(has=(rng.uniform(size=n_map) < 0.97).astype(np.uint8), in_kf=(rng.uniform(size=n_map) < 0.05).astype(np.uint8),
             nobs=rng.integers(1, 6, n_map).astype(np.int32))
    mbf = F32(40.0)
    uright = (np.where(rng.uniform(size=n_frame) < 0.5, d["keys"]["x"] - 5 + rng.normal(0, 1.5, n_frame), -1).astype(F32)
              if stereo else np.full(n_frame, -1, F32))
    bounds = (0.0, 0.0, 752.0, 480.0)
    R.set_bounds(bounds)
    kf = R.RefFrame(d["keys"], d["fdesc"], sf, uright=uright, mbf=mbf)
    R.set_camera(kf, *CAM)
    kf.set_pose(t)
    state = rng.choice(3, n_frame, p=[0.6, 0.3, 0.1])          # keyframe slot: empty / good point / bad point
    slot_obs = rng.integers(1, 6, n_frame).astype(np.int32)
    kf.set_mappoints(state > 0, nobs=slot_obs, bad=(state == 2))
    rn, ractions = R.fuse(kf, P, th)

    pc, u, v, invz, po, dist = _project_kf(P["xyz"], t)
    ok = _kf_valid(pc, u, v, po, dist, P, (0, 0, 752, 480)) & (P["has"] > 0) & (P["bad"] == 0) & (P["in_kf"] == 0)
    lvl = np.array([R.kf_predict_scale(kf, P["max_dist"][i], dist[i]) if ok[i] else 0 for i in range(n_map)], np.int32)
    pts = dict(u=u, v=v, ur=(u - (mbf * invz).astype(F32)).astype(F32), radius=(F32(th) * sf[lvl]).astype(F32),
               min_level=lvl - 1, max_level=lvl, valid=ok.astype(np.uint8), desc=P["desc"])
    inv_sigma2 = (F32(1.0) / (sf * sf)).astype(F32)
    bi, bd = O.search_window(d["keys"], d["fdesc"], uright, bounds, pts, 50, True, inv_sigma2)
    slot_mp = [("slot", i) if state[i] > 0 else None for i in range(n_frame)]
    n, actions, _ = _apply_fuse(bi, ok, slot_mp, state == 2, slot_obs, P["nobs"], sim3=False)
    assert rn == n and n > 200
    assert np.array_equal(ractions, actions)
    assert (actions[:, 0] == 1).sum() > 20 and (actions[:, 0] == 3).sum() > 20


def test_fuse_sim3_vs_reference():
    """ORBmatcher::Fuse(KeyFrame*, Sim3f& Scw, vpPoints, th, vpReplacePoint), ORBmatcher.cc:1536-1688."""
    n_map, n_frame, th = 3000, 1500, 4.0
    scw = (F32(1.25), F32(0.3), F32(-0.2), F32(0.55))
    t_eff = tuple(float(F32(c) / scw[0]) for c in scw[1:])
    d, rng, sf, P = _kf_case(n_map, n_frame, 7, t_eff)
    bounds = (0.0, 0.0, 752.0, 480.0)
    R.set_bounds(bounds)
    kf = R.RefFrame(d["keys"], d["fdesc"], sf)
    R.set_camera(kf, *CAM)
    state = rng.choice(3, n_frame, p=[0.6, 0.3, 0.1])
    kf.set_mappoints(state > 0, bad=(state == 2))
    rn, rrepl, ractions = R.fuse_sim3(kf, scw, P, th)

    pc, u, v, invz, po, dist = _project_kf(P["xyz"], scw[1:], with_sim3_scale=scw[0])
    ok = _kf_valid(pc, u, v, po, dist, P, (0, 0, 752, 480)) & (P["bad"] == 0)
    lvl = np.array([R.kf_predict_scale(kf, P["max_dist"][i], dist[i]) if ok[i] else 0 for i in range(n_map)], np.int32)
    pts = dict(u=u, v=v, radius=(F32(th) * sf[lvl]).astype(F32), min_level=lvl - 1, max_level=lvl, valid=ok.astype(np.uint8),
               desc=P["desc"])
    bi, bd = O.search_window(d["keys"], d["fdesc"], None, bounds, pts, 50, False, None)
    slot_mp = [("slot", i) if state[i] > 0 else None for i in range(n_frame)]
    n, actions, repl = _apply_fuse(bi, ok, slot_mp, state == 2, None, None, sim3=True)
    assert rn == n and n > 200
    # a candidate that lands on a slot filled earlier in the same call by another candidate: the reference stores
    # that candidate's pointer; the driver reports it as (id - 1000000), mirror it
    repl_ref_view = np.where(repl <= -100, (-100 - repl) - 1000000, repl)
    assert np.array_equal(rrepl, repl_ref_view) and np.array_equal(ractions, actions)
    assert (repl >= 0).sum() > 20


@pytest.mark.parametrize("ratio", [1.0, 1.5])
def test_search_by_projection_sim3_vs_reference(ratio):
    """ORBmatcher::SearchByProjection(KeyFrame*, Sim3f&, vpPoints, vpMatched, th, ratioHamming), ORBmatcher.cc:496-610."""
    n_map, n_frame, th = 4000, 1500, 8
    scw = (F32(0.8), F32(-0.4), F32(0.1), F32(0.3))
    t_eff = tuple(float(F32(c) / scw[0]) for c in scw[1:])
    d, rng, sf, P = _kf_case(n_map, n_frame, 9, t_eff)
    bounds = (0.0, 0.0, 752.0, 480.0)
    R.set_bounds(bounds)
    kf = R.RefFrame(d["keys"], d["fdesc"], sf)
    R.set_camera(kf, *CAM)
    matched = (rng.uniform(size=n_frame) < 0.2).astype(np.uint8)
    rn, rslots = R.search_kf_sim3(kf, scw, P, matched, th, ratio)

    pc, u, v, invz, po, dist = _project_kf(P["xyz"], scw[1:], with_sim3_scale=scw[0])
    ok = _kf_valid(pc, u, v, po, dist, P, (0, 0, 752, 480)) & (P["bad"] == 0)
    lvl = np.array([R.kf_predict_scale(kf, P["max_dist"][i], dist[i]) if ok[i] else 0 for i in range(n_map)], np.int32)
    pts = dict(u=u, v=v, ur=np.zeros(n_map, F32), radius=(F32(th) * sf[lvl]).astype(F32), min_level=lvl - 1, max_level=lvl,
               angle=np.zeros(n_map, F32), valid=ok.astype(np.uint8), blocks=np.ones(n_map, np.uint8), desc=P["desc"])
    th_acc = int(np.floor(F32(50) * F32(ratio)))
    n, asg, _, _ = O.search_by_projection(d["keys"], d["fdesc"], None, bounds, pts, 2, th_acc, 1.0, False, matched,
                                          np.full(n_frame, UNTOUCHED, np.int32), sf)
    assert rn == n and n > 200
    assert np.array_equal(rslots, asg)


@pytest.mark.parametrize("seed", [0, 1])
def test_search_by_sim3_vs_reference(seed):
    """ORBmatcher::SearchBySim3, ORBmatcher.cc:1690-1940."""
    from oracle.oracle import KP_DTYPE
    n1, n2, th = 1800, 2100, 7.5
    fx, fy, cx, cy = CAM
    rng = np.random.default_rng(seed + 40)
    d = synth.map_vs_frame(n1, n2, seed + 30, w=752, h=480)      # KF2 = the "frame", KF1's points = the "map"
    sf = d["scale_factors"]
    k2, d2 = d["keys"], d["fdesc"]
    k1 = np.zeros(n1, KP_DTYPE)
    k1["x"], k1["y"], k1["octave"] = d["u"], d["v"], d["level"]
    desc1 = d["mdesc"]
    t1w, t2w = np.array([0.1, 0.05, -0.2], F32), np.array([-0.3, 0.1, 0.25], F32)
    s12 = (F32(1.1), F32(0.35), F32(-0.1), F32(0.4))
    s, t12 = s12[0], np.array(s12[1:], F32)
    s21 = F32(1.0) / s
    t21 = np.array([-t12[0] / s, -t12[1] / s, -t12[2] / s], F32)

    def world_for(u, v, lvl, t_cam_from_world, sim_s, sim_t):
        """World points whose image in the OTHER keyframe (through the similarity) lands near (u, v)."""
        m = len(u)
        z = rng.uniform(2.0, 12.0, m); z[::47] = -1.0
        pc_other = np.stack([(u - float(cx)) * z / float(fx), (v - float(cy)) * z / float(fy), z], 1)
        pc_own = (pc_other - sim_t.astype(np.float64)) / float(sim_s)
        xyz = (pc_own - t_cam_from_world.astype(np.float64)).astype(F32)
        dist = np.linalg.norm(pc_other, axis=1)
        max_d = (dist * 1.2 ** (lvl - 0.5)).astype(F32); max_d[::29] *= F32(0.3)
        return xyz, (max_d / F32(sf[-1])).astype(F32), max_d

    def mirror(xyz, t_own, sim_s, sim_t):
        p_own = (xyz + t_own).astype(F32)
        p_oth = ((p_own * sim_s).astype(F32) + sim_t).astype(F32)
        with np.errstate(all="ignore"):
            invz = (1.0 / p_oth[:, 2].astype(np.float64)).astype(F32)
        x, y = (p_oth[:, 0] * invz).astype(F32), (p_oth[:, 1] * invz).astype(F32)
        u, v = ((fx * x).astype(F32) + cx).astype(F32), ((fy * y).astype(F32) + cy).astype(F32)
        dist = np.sqrt((p_oth[:, 0] * p_oth[:, 0] + p_oth[:, 1] * p_oth[:, 1]) + p_oth[:, 2] * p_oth[:, 2]).astype(F32)
        return p_oth, u, v, dist

    # KF1's map points -> KF2 through S21; KF2's map points -> KF1 through S12 (near their KF1 source keypoint, if any)
    xyz1, min1, max1 = world_for(d["u"].astype(np.float64), d["v"].astype(np.float64), d["level"], t1w, s21, t21)
    back = np.full(n2, -1)
    for i1 in range(n1):
        if d["src"][i1] >= 0:
            back[d["src"][i1]] = i1
    u21 = rng.uniform(0, 752, n2); v21 = rng.uniform(0, 480, n2)
    hasb = back >= 0
    u21[hasb] = k1["x"][back[hasb]] + rng.normal(0, 2, hasb.sum()); v21[hasb] = k1["y"][back[hasb]] + rng.normal(0, 2, hasb.sum())
    lvl2 = np.clip(k2["octave"] + (rng.uniform(size=n2) < 0.2), 0, 7).astype(np.int32)
    xyz2, min2, max2 = world_for(u21, v21, lvl2, t2w, s, t12)
    has1, has2 = rng.uniform(size=n1) < 0.9, rng.uniform(size=n2) < 0.9
    bad1, bad2 = rng.uniform(size=n1) < 0.03, rng.uniform(size=n2) < 0.03
    pre12 = np.full(n1, -1, np.int32)
    cand = np.flatnonzero(has2 & ~bad2)
    pre_idx = rng.choice(n1, 60, replace=False)
    pre12[pre_idx] = rng.choice(cand, 60, replace=False)

    bounds = (0.0, 0.0, 752.0, 480.0)
    R.set_bounds(bounds)
    kf1, kf2 = R.RefFrame(k1, desc1, sf), R.RefFrame(k2, d2, sf)
    for kf, t, has, bad, xyz, mn, mx, desc in ((kf1, t1w, has1, bad1, xyz1, min1, max1, desc1), (kf2, t2w, has2, bad2, xyz2, min2, max2, d2)):
        R.set_camera(kf, *CAM)
        kf.set_pose(t)
        kf.set_mappoints(has, xyz=xyz, desc=desc, bad=bad, min_dist=mn, max_dist=mx)
    rn, rout = R.search_by_sim3(kf1, kf2, s12, pre12, th)

    def side(xyz, t_own, sim_s, sim_t, has, bad, already, mn, mx, kf_other, desc):
        p, u, v, dist = mirror(xyz, t_own, sim_s, sim_t)
        ok = has & ~already & ~bad & ~(p[:, 2] < 0) & (u >= 0) & (u < 752) & (v >= 0) & (v < 480)
        ok &= ~((dist < F32(0.8) * mn) | (dist > F32(1.2) * mx))
        lvl = np.array([R.kf_predict_scale(kf_other, mx[i], dist[i]) if ok[i] else 0 for i in range(len(u))], np.int32)
        return dict(u=u, v=v, radius=(F32(th) * sf[lvl]).astype(F32), min_level=lvl - 1, max_level=lvl,
                    valid=ok.astype(np.uint8), desc=desc)
    already1 = pre12 >= 0
    already2 = np.zeros(n2, bool); already2[pre12[pre12 >= 0]] = True
    pts12 = side(xyz1, t1w, s21, t21, has1, bad1, already1, min1, max1, kf2, desc1)
    pts21 = side(xyz2, t2w, s, t12, has2, bad2, already2, min2, max2, kf1, d2)
    n, m12 = O.search_by_sim3(k1, desc1, k2, d2, bounds, pts12, pts21, 100)
    expect = np.where(m12 >= 0, m12, pre12)          # vpMatches12 keeps its earlier entries
    assert rn == n and n > 150
    assert np.array_equal(rout, expect)


def _distinctive_case(seed, n_points=300):
    """Map points observed by 1..40 keyframes (a few by 150); observations are noisy copies of a per-point descriptor,
    with outliers, so medians tie often."""
    rng = np.random.default_rng(seed)
    nkf = rng.integers(1, 41, n_points)
    nkf[::50] = 150
    nkf[1] = 1
    kf_start = np.concatenate([[0], np.cumsum(nkf)]).astype(np.int32)
    rows = rng.choice([1, 2], kf_start[-1], p=[0.8, 0.2]).astype(np.int32)
    kf_bad = (rng.uniform(size=kf_start[-1]) < 0.1).astype(np.uint8)
    kf_bad[kf_start[5]:kf_start[6]] = 1                                # a point whose keyframes are all bad
    descs = []
    for p in range(n_points):
        base = rng.integers(0, 256, 32).astype(np.uint8)
        for k in range(kf_start[p], kf_start[p + 1]):
            for _ in range(rows[k]):
                noise = 256 if rng.uniform() < 0.1 else int(rng.integers(0, 40))
                descs.append(synth.flip_bits(base, min(noise, 256), rng) if noise < 256 else rng.integers(0, 256, 32).astype(np.uint8))
    return np.stack(descs), kf_start, rows, kf_bad


def flatten_good_observations(desc, kf_start, rows, kf_bad):
    """What the caller hands to the batched API: the rows of the good keyframes, in vDescriptors order."""
    row_start = np.concatenate([[0], np.cumsum(rows)])
    keep, start = [], [0]
    for p in range(len(kf_start) - 1):
        for k in range(kf_start[p], kf_start[p + 1]):
            if not kf_bad[k]:
                keep.extend(range(row_start[k], row_start[k + 1]))
        start.append(len(keep))
    return desc[keep], np.array(start, np.int32)


@pytest.mark.parametrize("seed", [0, 1])
def test_distinctive_descriptors_vs_reference(seed):
    """MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:438-529."""
    desc, kf_start, rows, kf_bad = _distinctive_case(seed)
    ref = R.distinctive(desc, kf_start, rows, kf_bad)
    good, start = flatten_good_observations(desc, kf_start, rows, kf_bad)
    best = O.distinctive_descriptors(good, start)
    assert best[5] == -1 and (best >= 0).sum() >= len(best) - 3
    for p in range(len(best)):
        if best[p] >= 0:
            assert np.array_equal(good[start[p] + best[p]], ref[p]), p
        else:
            assert not ref[p].any()
