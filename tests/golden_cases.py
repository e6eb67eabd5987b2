"""Matcher / BoW golden cases shared by tests/golden/make_golden_matchers.py (runs the REFERENCE's own functions from
oracle/_ref/libref_orbmatcher.so and stores their outputs) and tests/test_golden.py (re-creates the same seeded inputs
on the GPU box, runs the CUDA entry points and compares with the stored reference outputs -- no oracle in between).

Every case is a class with
    ref()        -> dict of reference outputs (needs oracle/_ref; build container only)
    cuda(orbfe, stored) -> dict of the same keys computed by the CUDA library
`stored` carries the reference-side derived inputs that cannot be recomputed without the reference (PredictScale
levels, the fundamental matrix the reference built)."""
import hashlib

import numpy as np

import synth
from test_oracle_match_vs_ref import (CAM, F32, UNTOUCHED, _apply_fuse, _distinctive_case, _kf_case, _kf_valid,
                                       _lastframe_case, _mappoint_case, _project_kf, flatten_good_observations)
from test_oracle_bow_vs_ref import SF, _bow_frames, _tri_case


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest()[:8], np.uint8).copy()


def _norm(asg, initial_null):
    o = asg.copy()
    o[(o == -1) & initial_null] = UNTOUCHED
    return o


class MapPoints:
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, ...)"""
    name = "search_mappoints"

    def __init__(self):
        self.d, rng, self.mp, self.pts = _mappoint_case(3000, 800, 1, 3.0, False)
        n = 800
        self.uright = np.where(rng.uniform(size=n) < 0.5, self.d["keys"]["x"] - 5 + rng.normal(0, 3, n), -1).astype(np.float32)
        self.state = rng.choice(3, n, p=[0.9, 0.06, 0.04])

    def inputs(self):
        return digest(self.d["keys"], self.d["fdesc"], self.pts["u"], self.pts["desc"], self.uright, self.state)

    def ref(self):
        from oracle import ref as R
        R.set_bounds(self.d["bounds"])
        F = R.RefFrame(self.d["keys"], self.d["fdesc"], self.d["scale_factors"], uright=self.uright)
        F.set_mappoints(self.state > 0, nobs=(self.state == 1).astype(np.int32))
        n, slots = F.search_mappoints(self.mp, 3.0, False, 40.0, 0.8)
        return dict(n=np.int32(n), slots=slots)

    def cuda(self, orbfe, stored):
        F = orbfe.FrameData(self.d["keys"], self.d["fdesc"], self.d["bounds"], self.uright)
        n, asg, _, _ = orbfe.ORBmatcher(0.8, False).SearchByProjection(F, self.pts, (self.state == 1).astype(np.uint8),
                                                                         np.full(800, UNTOUCHED, np.int32))
        return dict(n=np.int32(n), slots=_norm(asg, self.state == 0))


class LastFrame:
    """ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono), camera moving forward"""
    name = "search_lastframe"

    def __init__(self):
        self.d, rng, self.last, self.pts = _lastframe_case(3000, 900, 4, "forward", False)
        n = 900
        self.uright = np.where(rng.uniform(size=n) < 0.5, self.d["keys"]["x"] - 20 + rng.normal(0, 8, n), -1).astype(np.float32)
        self.state = rng.choice(3, n, p=[0.9, 0.06, 0.04])

    def inputs(self):
        return digest(self.d["keys"], self.d["fdesc"], self.pts["u"], self.pts["desc"], self.uright, self.state, self.last["xyz"])

    def ref(self):
        from oracle import ref as R
        d, last = self.d, self.last
        R.set_bounds(d["bounds"])
        cur = R.RefFrame(d["keys"], d["fdesc"], d["scale_factors"], uright=self.uright, mb=last["mb"], mbf=last["mbf"])
        cur.set_mappoints(self.state > 0, nobs=(self.state == 1).astype(np.int32))
        lf = R.RefFrame(last["keys"], d["mdesc"], d["scale_factors"])
        lf.set_mappoints(last["has"], nobs=last["nobs"], xyz=last["xyz"], desc=d["mdesc"], outlier=last["outlier"])
        lf.set_pose(last["pose"])
        n, slots = cur.search_lastframe(lf, last["th"], False, 0.9, True)
        return dict(n=np.int32(n), slots=slots)

    def cuda(self, orbfe, stored):
        F = orbfe.FrameData(self.d["keys"], self.d["fdesc"], self.d["bounds"], self.uright)
        n, asg, _, _ = orbfe.ORBmatcher(0.9, True).SearchByProjectionLastFrame(F, self.pts, (self.state == 1).astype(np.uint8),
                                                                                 np.full(900, UNTOUCHED, np.int32))
        return dict(n=np.int32(n), slots=_norm(asg, self.state == 0))


class Fuse:
    """ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>&, th, bRight=false) incl. the graph updates it performs"""
    name = "fuse"

    def __init__(self):
        n_map, n_frame = 3000, 1500
        self.t = (0.21, -0.13, 0.4)
        self.d, rng, self.sf, self.P = _kf_case(n_map, n_frame, 1, self.t)
        self.P.update(has=(rng.uniform(size=n_map) < 0.97).astype(np.uint8), in_kf=(rng.uniform(size=n_map) < 0.05).astype(np.uint8),
                      nobs=rng.integers(1, 6, n_map).astype(np.int32))
        self.mbf = F32(40.0)
        self.uright = np.where(rng.uniform(size=n_frame) < 0.5, self.d["keys"]["x"] - 5 + rng.normal(0, 1.5, n_frame), -1).astype(F32)
        self.state = rng.choice(3, n_frame, p=[0.6, 0.3, 0.1])
        self.slot_obs = rng.integers(1, 6, n_frame).astype(np.int32)
        self.bounds = (0.0, 0.0, 752.0, 480.0)
        pc, self.u, self.v, self.invz, po, self.dist = _project_kf(self.P["xyz"], self.t)
        self.ok = _kf_valid(pc, self.u, self.v, po, self.dist, self.P, (0, 0, 752, 480)) & (self.P["has"] > 0) & \
            (self.P["bad"] == 0) & (self.P["in_kf"] == 0)

    def inputs(self):
        return digest(self.d["keys"], self.d["fdesc"], self.P["xyz"], self.P["desc"], self.uright, self.state, self.u, self.v)

    def _kf(self):
        from oracle import ref as R
        R.set_bounds(self.bounds)
        kf = R.RefFrame(self.d["keys"], self.d["fdesc"], self.sf, uright=self.uright, mbf=self.mbf)
        R.set_camera(kf, *CAM)
        kf.set_pose(self.t)
        kf.set_mappoints(self.state > 0, nobs=self.slot_obs, bad=(self.state == 2))
        return kf

    def ref(self):
        from oracle import ref as R
        kf = self._kf()
        n, actions = R.fuse(kf, self.P, 3.0)
        lvl = np.array([R.kf_predict_scale(kf, self.P["max_dist"][i], self.dist[i]) if self.ok[i] else 0
                        for i in range(len(self.ok))], np.int8)
        return dict(n=np.int32(n), actions=actions, lvl=lvl)

    def cuda(self, orbfe, stored):
        lvl = stored["lvl"].astype(np.int32)
        pts = dict(u=self.u, v=self.v, ur=(self.u - (self.mbf * self.invz).astype(F32)).astype(F32),
                   radius=(F32(3.0) * self.sf[lvl]).astype(F32), min_level=lvl - 1, max_level=lvl,
                   valid=self.ok.astype(np.uint8), desc=self.P["desc"])
        F = orbfe.FrameData(self.d["keys"], self.d["fdesc"], self.bounds, self.uright)
        _, bi, _ = orbfe.ORBmatcher().FuseSearch(F, pts, (F32(1.0) / (self.sf * self.sf)).astype(F32))
        slot_mp = [("slot", i) if self.state[i] > 0 else None for i in range(len(self.state))]
        n, actions, _ = _apply_fuse(bi, self.ok, slot_mp, self.state == 2, self.slot_obs, self.P["nobs"], sim3=False)
        return dict(n=np.int32(n), actions=actions, lvl=stored["lvl"])


class Bow:
    """DBoW2 transform of two frames and ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)"""
    name = "bow"

    def __init__(self):
        self.voc = synth.make_vocabulary(10, 4, 3)
        self.kk, self.dk, self.kf_, self.df, rng = _bow_frames(self.voc, 1500, 1400, 1)
        self.state = rng.choice(3, len(self.kk), p=[0.2, 0.7, 0.1])

    def inputs(self):
        return digest(self.voc["desc"], self.voc["weight"], self.dk, self.df, self.kk["angle"], self.kf_["angle"], self.state)

    def ref(self):
        import os
        import tempfile
        from oracle import ref as R
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "voc.txt")
            synth.write_vocabulary_text(path, self.voc)
            rv = R.RefVocabulary(path)
        word, weight, node = rv.transform_features(self.dk, 2)
        (ids, vals), _ = rv.transform(self.dk, 2)
        R.set_bounds((0.0, 0.0, 752.0, 480.0))
        KF, F = R.RefFrame(self.kk, self.dk, SF), R.RefFrame(self.kf_, self.df, SF)
        KF.set_mappoints(self.state > 0, bad=(self.state == 2))
        R.compute_bow(KF, rv, 2); R.compute_bow(F, rv, 2)
        n, out = R.search_by_bow_kf_f(KF, F, 0.7, True)
        return dict(word=word, node=node, bow_ids=ids.astype(np.int64), bow_vals=vals, n=np.int32(n), matches=out)

    def cuda(self, orbfe, stored):
        gv = orbfe.ORBVocabulary(10, 4, self.voc["parent"], self.voc["desc"], self.voc["weight"])
        word, weight, node = gv.transform_features(self.dk, 2)
        (ids, vals), fva = gv.transform(self.dk, 2)
        _, fvb = gv.transform(self.df, 2)
        n, mA, _ = orbfe.ORBmatcher(0.7, True).SearchByBoW((self.dk, self.kk["angle"], self.state == 1, fva),
                                                            (self.df, self.kf_["angle"], None, fvb))
        out = np.full(len(self.kf_), -1, np.int32)
        out[mA[mA >= 0]] = np.flatnonzero(mA >= 0)
        return dict(word=word, node=node, bow_ids=ids.astype(np.int64), bow_vals=vals, n=np.int32(n), matches=out)


class Triangulation:
    """ORBmatcher::SearchForTriangulation + Pinhole::epipolarConstrain"""
    name = "triangulation"

    def __init__(self):
        self.voc = synth.make_vocabulary(10, 4, 3)
        self.c = _tri_case(self.voc, 2, 0.5)

    def inputs(self):
        k1, d1, ur1, mp1, k2, d2, ur2, mp2 = self.c
        return digest(k1, d1, ur1, mp1, k2, d2, ur2, mp2)

    def ref(self):
        import os
        import tempfile
        from oracle import ref as R
        k1, d1, ur1, mp1, k2, d2, ur2, mp2 = self.c
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "voc.txt")
            synth.write_vocabulary_text(path, self.voc)
            rv = R.RefVocabulary(path)
        R.set_bounds((0.0, 0.0, 752.0, 480.0))
        KF1, KF2 = R.RefFrame(k1, d1, SF, uright=ur1), R.RefFrame(k2, d2, SF, uright=ur2)
        for kf, t in ((KF1, (0.0, 0.0, 0.0)), (KF2, (-0.3, 0.01, 0.02))):
            R.set_camera(kf, 458.654, 457.296, 367.215, 248.375)
            kf.set_pose(t)
        KF1.set_mappoints(mp1); KF2.set_mappoints(mp2)
        R.compute_bow(KF1, rv, 2); R.compute_bow(KF2, rv, 2)
        n, m12, f12, ep = R.search_for_triangulation(KF1, KF2, False, False, False)
        return dict(n=np.int32(n), m12=m12, f12=f12, ep=ep)

    def cuda(self, orbfe, stored):
        k1, d1, ur1, mp1, k2, d2, ur2, mp2 = self.c
        gv = orbfe.ORBVocabulary(10, 4, self.voc["parent"], self.voc["desc"], self.voc["weight"])
        _, fva = gv.transform(d1, 2)
        _, fvb = gv.transform(d2, 2)
        n, m12 = orbfe.ORBmatcher(0.6, False).SearchForTriangulation((k1, d1, ur1, mp1, fva), (k2, d2, ur2, mp2, fvb),
                                                                      stored["f12"], stored["ep"], SF, SF * SF, False, False)
        return dict(n=np.int32(n), m12=m12, f12=stored["f12"], ep=stored["ep"])


class Stereo:
    """ORBextractor on both images + Frame::ComputeStereoMatches"""
    name = "stereo"

    def __init__(self):
        self.L, self.R = synth.stereo_pair(480, 752, 0)

    def inputs(self):
        return digest(self.L, self.R)

    def ref(self):
        from oracle import ref as R
        kl, dl, kr, dr, ur, dp = R.stereo(self.L, self.R, 40.0, 0.11)
        return dict(nl=np.int32(len(kl)), nr=np.int32(len(kr)), kl=digest(kl), dl=digest(dl), uright=ur, depth=dp)

    def cuda(self, orbfe, stored):
        gl, gr = orbfe.ORBextractor(1000), orbfe.ORBextractor(1000)
        _, kl, dl = gl(self.L, None, (0, 0))
        _, kr, dr = gr(self.R, None, (0, 0))
        ur, dp = orbfe.ORBmatcher.ComputeStereoMatches(gl, gr, kl, dl, kr, dr, 40.0, 0.11)
        return dict(nl=np.int32(len(kl)), nr=np.int32(len(kr)), kl=digest(kl), dl=digest(dl), uright=ur, depth=dp)


class Distinctive:
    """MapPoint::ComputeDistinctiveDescriptors for 300 map points"""
    name = "distinctive"

    def __init__(self):
        self.desc, self.kf_start, self.rows, self.kf_bad = _distinctive_case(0)

    def inputs(self):
        return digest(self.desc, self.kf_start, self.rows, self.kf_bad)

    def ref(self):
        from oracle import ref as R
        return dict(chosen=R.distinctive(self.desc, self.kf_start, self.rows, self.kf_bad))

    def cuda(self, orbfe, stored):
        good, start = flatten_good_observations(self.desc, self.kf_start, self.rows, self.kf_bad)
        best = orbfe.ORBmatcher.ComputeDistinctiveDescriptors(good, start)
        out = np.zeros((len(best), 32), np.uint8)
        for p in range(len(best)):
            if best[p] >= 0:
                out[p] = good[start[p] + best[p]]
        return dict(chosen=out)


CASES = [MapPoints, LastFrame, Fuse, Bow, Triangulation, Stereo, Distinctive]
