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
