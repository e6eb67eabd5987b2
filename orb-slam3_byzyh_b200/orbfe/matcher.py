"""ORBmatcher: the Frame-based hot-path subset of ORB_SLAM3::ORBmatcher
(/root/reference/include/ORBmatcher.h:33-104, src/ORBmatcher.cc) plus the two stereo associators
of ORB_SLAM3::Frame (src/Frame.cc:1102-1358, 1530-1587), over the C ABI."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._lib import BowSide, TriParams, TriSide, FrameView, ProjPoints, SearchParams, WindowParams, check, lib, ptr

MODE_MAPPOINTS, MODE_LASTFRAME, MODE_KEYFRAME = 0, 1, 2


@dataclass
class FrameData:
    """The part of ORB_SLAM3::Frame the matchers read (include/Frame.h; Nleft == -1 layout)."""
    keys: np.ndarray                    # mvKeysUn, KP_DTYPE[N]
    desc: np.ndarray                    # mDescriptors, uint8[N,32]
    bounds: tuple                       # (mnMinX, mnMinY, mnMaxX, mnMaxY)
    uright: Optional[np.ndarray] = None  # mvuRight

    def view(self, keep):
        keys = np.ascontiguousarray(self.keys)
        desc = np.ascontiguousarray(self.desc, np.uint8)
        keep += [keys, desc]
        fv = FrameView()
        fv.n = len(keys)
        fv.keys, fv.desc = keys.ctypes.data, desc.ctypes.data
        if self.uright is not None:
            ur = np.ascontiguousarray(self.uright, np.float32)
            keep.append(ur)
            fv.uright = ur.ctypes.data
        b = [np.float32(v) for v in self.bounds]
        fv.min_x, fv.min_y, fv.max_x, fv.max_y = b
        # mfGridElementWidthInv / HeightInv, src/Frame.cc:303-305 (FRAME_GRID_COLS=64, ROWS=48)
        fv.grid_w_inv = np.float32(64) / np.float32(b[2] - b[0])
        fv.grid_h_inv = np.float32(48) / np.float32(b[3] - b[1])
        return fv


class ORBmatcher:
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30   # ORBmatcher.cc:36-38

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio, self.mbCheckOrientation, self.device = float(nnratio), bool(checkOri), device

    # static int DescriptorDistance(a, b), ORBmatcher.cc:2384-2404 (batched over rows)
    @staticmethod
    def DescriptorDistance(a, b, device=0):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        assert a.shape == b.shape
        out = np.empty(len(a), np.int32)
        check(lib().orbfe_descriptor_distance(ptr(a), ptr(b), len(a), ptr(out), device))
        return int(out[0]) if len(out) == 1 else out

    def _search(self, F, pts, mode, th_accept, claimed, assigned):
        keep = []
        fv = F.view(keep)
        pp = ProjPoints()
        m = len(pts["u"])
        pp.m = m
        for name, dt in [("u", np.float32), ("v", np.float32), ("ur", np.float32), ("radius", np.float32),
                         ("min_level", np.int32), ("max_level", np.int32), ("angle", np.float32),
                         ("valid", np.uint8), ("blocks", np.uint8), ("desc", np.uint8)]:
            if pts.get(name) is None:
                continue
            a = np.ascontiguousarray(pts[name], dt)
            keep.append(a)
            setattr(pp, name, a.ctypes.data)
        prm = SearchParams(mode, th_accept, self.mfNNratio, int(self.mbCheckOrientation))
        claimed = np.ascontiguousarray(claimed, np.uint8)
        assigned = np.ascontiguousarray(assigned, np.int32).copy()
        bi, bd = np.empty(m, np.int32), np.empty(m, np.int32)
        n = check(lib().orbfe_search_by_projection(C.byref(fv), C.byref(pp), C.byref(prm), ptr(claimed),
                                                   ptr(assigned), ptr(bi), ptr(bd), self.device))
        return n, assigned, bi, bd

    # SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints), ORBmatcher.cc:46
    def SearchByProjection(self, F, pts, claimed, assigned):
        return self._search(F, pts, MODE_MAPPOINTS, self.TH_HIGH, claimed, assigned)

    # SearchByProjection(Frame& Cur, const Frame& Last, th, bMono), ORBmatcher.cc:1951
    def SearchByProjectionLastFrame(self, F, pts, claimed, assigned):
        return self._search(F, pts, MODE_LASTFRAME, self.TH_HIGH, claimed, assigned)

    # SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, th, ORBdist), ORBmatcher.cc:2197
    def SearchByProjectionKeyFrame(self, F, pts, claimed, assigned, ORBdist):
        return self._search(F, pts, MODE_KEYFRAME, int(ORBdist), claimed, assigned)

    # The map-point (mode 0) and last-frame (mode 1) overloads on a fisheye stereo frame (Nleft != -1):
    # ORBmatcher.cc:46-240 incl. :171-237, and :1951-2185 incl. :2090-2155
    def SearchByProjectionFisheye(self, FL, FR, l2r, r2l, ptsL, ptsR, claimed, assigned, mode=MODE_MAPPOINTS):
        keep = []
        fl, fr = FL.view(keep), FR.view(keep)

        def proj(d):
            pp = ProjPoints()
            pp.m = len(d["u"])
            for name, dt in [("u", np.float32), ("v", np.float32), ("ur", np.float32), ("radius", np.float32),
                             ("min_level", np.int32), ("max_level", np.int32), ("angle", np.float32),
                             ("valid", np.uint8), ("blocks", np.uint8), ("desc", np.uint8)]:
                if d.get(name) is None:
                    continue
                a = np.ascontiguousarray(d[name], dt)
                keep.append(a)
                setattr(pp, name, a.ctypes.data)
            return pp
        pl, pr = proj(ptsL), proj(ptsR)
        prm = SearchParams(mode, self.TH_HIGH, self.mfNNratio, int(self.mbCheckOrientation))
        l2r = np.ascontiguousarray(l2r, np.int32)
        r2l = np.ascontiguousarray(r2l, np.int32)
        claimed = np.ascontiguousarray(claimed, np.uint8)
        assigned = np.ascontiguousarray(assigned, np.int32).copy()
        bl, br = np.empty(pl.m, np.int32), np.empty(pl.m, np.int32)
        n = check(lib().orbfe_search_by_projection_fisheye(C.byref(fl), C.byref(fr), ptr(l2r), ptr(r2l), C.byref(pl),
                                                           C.byref(pr), C.byref(prm), ptr(claimed), ptr(assigned),
                                                           ptr(bl), ptr(br), self.device))
        return n, assigned, bl, br

    # SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:735-891
    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=100):
        keep = []
        f1, f2 = F1.view(keep), F2.view(keep)
        prev = np.ascontiguousarray(vbPrevMatched, np.float32).copy()
        m12 = np.empty(f1.n, np.int32)
        n = check(lib().orbfe_search_for_initialization(C.byref(f1), C.byref(f2), ptr(prev), int(windowSize),
                                                        self.mfNNratio, int(self.mbCheckOrientation), ptr(m12),
                                                        self.device))
        return n, m12, prev

    @staticmethod
    def _proj(pts, keep):
        pp = ProjPoints()
        pp.m = len(pts["u"])
        for name, dt in [("u", np.float32), ("v", np.float32), ("ur", np.float32), ("radius", np.float32),
                         ("min_level", np.int32), ("max_level", np.int32), ("angle", np.float32),
                         ("valid", np.uint8), ("blocks", np.uint8), ("desc", np.uint8)]:
            if pts.get(name) is None:
                continue
            a = np.ascontiguousarray(pts[name], dt)
            keep.append(a)
            setattr(pp, name, a.ctypes.data)
        return pp

    # Inner loop of Fuse(KeyFrame*, vector<MapPoint*>&, th, bRight) (ORBmatcher.cc:1326-1534; pass
    # inv_level_sigma2 = pKF->mvInvLevelSigma2 for its reprojection gate) and of Fuse(KeyFrame*, Sim3f&, ...)
    # (:1536-1688, no gate): per point the keypoint it fuses with (or -1) and the best distance.
    def FuseSearch(self, KF, pts, inv_level_sigma2=None, th=None):
        keep = []
        fv = KF.view(keep)
        pp = self._proj(pts, keep)
        prm = WindowParams(self.TH_LOW if th is None else int(th), 0, None, 0)
        if inv_level_sigma2 is not None:
            inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
            keep.append(inv)
            prm.gate, prm.inv_level_sigma2, prm.n_levels = 1, inv.ctypes.data, len(inv)
        bi, bd = np.empty(pp.m, np.int32), np.empty(pp.m, np.int32)
        n = check(lib().orbfe_search_window(C.byref(fv), C.byref(pp), C.byref(prm), ptr(bi), ptr(bd), self.device))
        return n, bi, bd

    # SearchBySim3(pKF1, pKF2, vpMatches12, S12, th), ORBmatcher.cc:1690-1940
    def SearchBySim3(self, KF1, KF2, pts12, pts21):
        keep = []
        f1, f2 = KF1.view(keep), KF2.view(keep)
        p12, p21 = self._proj(pts12, keep), self._proj(pts21, keep)
        m12 = np.empty(f1.n, np.int32)
        n = check(lib().orbfe_search_by_sim3(C.byref(f1), C.byref(f2), C.byref(p12), C.byref(p21), self.TH_HIGH,
                                             ptr(m12), self.device))
        return n, m12

    # SearchByProjection(KeyFrame*, Sim3f& Scw, vpPoints, vpMatched, th, ratioHamming), ORBmatcher.cc:496-733
    def SearchByProjectionSim3(self, KF, pts, matched, assigned, ratioHamming=1.0):
        keep_ori = self.mbCheckOrientation
        self.mbCheckOrientation = False
        try:
            th = int(np.floor(np.float32(self.TH_LOW) * np.float32(ratioHamming)))
            return self._search(KF, pts, MODE_KEYFRAME, th, matched, assigned)
        finally:
            self.mbCheckOrientation = keep_ori

    def _bow(self, sideA, sideB, strict, n_left_b):
        """side = (desc, angle, valid or None, (nodes, start, feat))."""
        keep = []

        def mk(side):
            desc, angle, valid, fv = side
            s = BowSide()
            desc = np.ascontiguousarray(desc, np.uint8)
            angle = np.ascontiguousarray(angle, np.float32)
            nodes, start, feat = [np.ascontiguousarray(x, np.int32) for x in fv]
            keep.extend([desc, angle, nodes, start, feat])
            s.n, s.desc, s.angle = len(desc), desc.ctypes.data, angle.ctypes.data
            if valid is not None:
                valid = np.ascontiguousarray(valid, np.uint8)
                keep.append(valid)
                s.valid = valid.ctypes.data
            s.fv.n_nodes, s.fv.node_id, s.fv.start, s.fv.feat = len(nodes), nodes.ctypes.data, start.ctypes.data, feat.ctypes.data
            return s
        a, b = mk(sideA), mk(sideB)
        mA = np.empty(a.n, np.int32)
        mR = np.empty(a.n, np.int32)
        n = check(lib().orbfe_search_by_bow(C.byref(a), C.byref(b), self.TH_LOW, int(strict), self.mfNNratio,
                                            int(self.mbCheckOrientation), int(n_left_b), ptr(mA), ptr(mR), self.device))
        return n, mA, mR

    # SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches), ORBmatcher.cc:260-494.  kf / f = (desc, angles,
    # valid, feature vector); valid of the keyframe = "slot holds a good map point", of the frame None.
    # -> nmatches, match[iKF] = frame feature (left camera), matchR[iKF] = right camera (fisheye frames only).
    def SearchByBoW(self, kf, f, n_left=-1):
        return self._bow(kf, f, False, n_left)

    # SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12), ORBmatcher.cc:893-1044
    def SearchByBoWKeyFrames(self, kf1, kf2):
        n, mA, _ = self._bow(kf1, kf2, True, -1)
        return n, mA

    # SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), ORBmatcher.cc:1046-1324 (pinhole).
    # kf = (keys, desc, uright or None, has_map_point, feature vector); f12 = 3x3 fundamental matrix of
    # Pinhole::epipolarConstrain, epipole = project(T2w * Cw).  -> nmatches, matches12
    def SearchForTriangulation(self, kf1, kf2, f12, epipole, scale_factors2, level_sigma2_2, bOnlyStereo=False, bCoarse=False,
                               rig=None):
        """rig (two-camera keyframes, ORBmatcher.cc:1071-1095, 1160-1241): dict(n_left1, n_left2, level_sigma2_1,
        pairs=[(params1[8], params2[8], precision1, precision2, R12 3x3, t12 3)] x 4 for ll, lr, rl, rr); the sides'
        keys / desc are then [mvKeys | mvKeysRight] and uright is None."""
        keep = []

        def mk(side):
            keys, desc, uright, has_mp, fv = side
            t = TriSide()
            keys = np.ascontiguousarray(keys)
            desc = np.ascontiguousarray(desc, np.uint8)
            has_mp = np.ascontiguousarray(has_mp, np.uint8)
            nodes, start, feat = [np.ascontiguousarray(x, np.int32) for x in fv]
            keep.extend([keys, desc, has_mp, nodes, start, feat])
            t.n, t.keys, t.desc, t.has_map_point = len(keys), keys.ctypes.data, desc.ctypes.data, has_mp.ctypes.data
            if uright is not None:
                uright = np.ascontiguousarray(uright, np.float32)
                keep.append(uright)
                t.uright = uright.ctypes.data
            t.fv.n_nodes, t.fv.node_id, t.fv.start, t.fv.feat = len(nodes), nodes.ctypes.data, start.ctypes.data, feat.ctypes.data
            return t
        a, b = mk(kf1), mk(kf2)
        sf = np.ascontiguousarray(scale_factors2, np.float32)
        s2 = np.ascontiguousarray(level_sigma2_2, np.float32)
        prm = TriParams()
        prm.f12 = (C.c_float * 9)(*[float(x) for x in np.asarray(f12, np.float32).reshape(-1)])
        prm.epipole = (C.c_float * 2)(*[float(x) for x in np.asarray(epipole, np.float32).reshape(-1)])
        prm.scale_factors2, prm.level_sigma2_2, prm.n_levels = sf.ctypes.data, s2.ctypes.data, len(sf)
        prm.only_stereo, prm.coarse = int(bOnlyStereo), int(bCoarse)
        prm.check_orientation, prm.th_low = int(self.mbCheckOrientation), self.TH_LOW
        if rig is not None:
            from ._lib import TriRig
            r = TriRig()
            r.n_left1, r.n_left2 = int(rig["n_left1"]), int(rig["n_left2"])
            s1 = np.ascontiguousarray(rig["level_sigma2_1"], np.float32)
            keep.append(s1)
            r.level_sigma2_1 = s1.ctypes.data
            for k, (p1, p2, pr1, pr2, R12, t12) in enumerate(rig["pairs"]):
                c = r.pair[k]
                c.params1 = (C.c_float * 8)(*[float(x) for x in p1])
                c.params2 = (C.c_float * 8)(*[float(x) for x in p2])
                c.precision1, c.precision2 = float(pr1), float(pr2)
                c.R12 = (C.c_float * 9)(*[float(x) for x in np.asarray(R12, np.float32).reshape(-1)])
                c.t12 = (C.c_float * 3)(*[float(x) for x in np.asarray(t12, np.float32).reshape(-1)])
            keep.append(r)
            prm.rig = C.pointer(r)
        m12 = np.empty(a.n, np.int32)
        n = check(lib().orbfe_search_for_triangulation(C.byref(a), C.byref(b), C.byref(prm), ptr(m12), self.device))
        return n, m12

    # MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:438-529) for a batch of map points: desc = all observed
    # descriptors, start[p] .. start[p+1] = the rows of point p.  -> best row per point (relative), -1 = none
    @staticmethod
    def ComputeDistinctiveDescriptors(desc, start, device=0):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        start = np.ascontiguousarray(start, np.int32)
        best = np.empty(len(start) - 1, np.int32)
        check(lib().orbfe_distinctive_descriptors(ptr(desc), ptr(start), len(start) - 1, ptr(best), device))
        return best

    # cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) + 0.7 ratio, Frame.cc:1553-1562
    def knn2(self, query, train, train_offset=0):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        idx, dist = np.empty((len(q), 2), np.int32), np.empty((len(q), 2), np.int32)
        match = np.empty(len(q), np.int32)
        check(lib().orbfe_knn2(ptr(q), len(q), ptr(t), len(t), train_offset, ptr(idx), ptr(dist), ptr(match),
                               self.device))
        return idx, dist, match

    def knn2_merge(self, idx_shards, dist_shards):
        idx_shards = np.ascontiguousarray(idx_shards, np.int32)
        dist_shards = np.ascontiguousarray(dist_shards, np.int32)
        G, nq = idx_shards.shape[:2]
        idx, dist = np.empty((nq, 2), np.int32), np.empty((nq, 2), np.int32)
        match = np.empty(nq, np.int32)
        check(lib().orbfe_knn2_merge(ptr(idx_shards), ptr(dist_shards), G, nq, ptr(idx), ptr(dist), ptr(match),
                                     self.device))
        return idx, dist, match

    # Frame::ComputeStereoMatches, Frame.cc:1102-1358
    @staticmethod
    def ComputeStereoMatches(exL, exR, keysL, descL, keysR, descR, mbf, mb, frame=0):
        keysL, keysR = np.ascontiguousarray(keysL), np.ascontiguousarray(keysR)
        descL = np.ascontiguousarray(descL, np.uint8)
        descR = np.ascontiguousarray(descR, np.uint8)
        ur, dp = np.empty(len(keysL), np.float32), np.empty(len(keysL), np.float32)
        check(lib().orbfe_stereo_match(exL.handle, exR.handle, frame, ptr(keysL), ptr(descL), len(keysL), ptr(keysR),
                                       ptr(descR), len(keysR), mbf, mb, ptr(ur), ptr(dp)))
        return ur, dp

    # The same for a batch of rectified pairs resident in HBM (device tensors: the output slabs of two
    # ORBextractor.extract_batch_device calls on the same stream) -> (uRight, depth) [B, capacity] float32 CUDA tensors
    @staticmethod
    def ComputeStereoMatchesBatchDevice(exL, exR, d_kpsL, d_descL, d_nL, d_kpsR, d_descR, d_nR, mbf, mb, stream=None):
        import torch
        B, cap = d_kpsL.shape[0], d_kpsL.shape[1]
        ur = torch.empty((B, cap), dtype=torch.float32, device=d_kpsL.device)
        dp = torch.empty_like(ur)
        st = C.c_void_p(stream.cuda_stream) if stream is not None else None
        check(lib().orbfe_stereo_match_batch_device(exL.handle, exR.handle, B, ptr(d_kpsL), ptr(d_descL), ptr(d_nL), ptr(d_kpsR),
                                                    ptr(d_descR), ptr(d_nR), cap, mbf, mb, ptr(ur), ptr(dp), st))
        return ur, dp

    # Frame::ComputeStereoFishEyeMatches' kNN-2 + ratio test (Frame.cc:1545-1562) for a batch of pairs resident in HBM:
    # queries = left rows [monoL[b], nL[b]), train = right rows [monoR[b], nR[b]) -> idx2, dist2 [B, cap, 2], match [B, cap]
    @staticmethod
    def knn2_batch_device(d_descL, d_monoL, d_nL, d_descR, d_monoR, d_nR, stream=None):
        import torch
        B, cap = d_descL.shape[0], d_descL.shape[1]
        with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(d_descL.device)):
            # the fills are ordered before the kernel: same stream
            idx2 = torch.full((B, cap, 2), -1, dtype=torch.int32, device=d_descL.device)
            dist2 = torch.full((B, cap, 2), -1, dtype=torch.int32, device=d_descL.device)
            match = torch.full((B, cap), -1, dtype=torch.int32, device=d_descL.device)
        st = C.c_void_p(stream.cuda_stream) if stream is not None else None
        check(lib().orbfe_knn2_batch_device(ptr(d_descL), ptr(d_monoL), ptr(d_nL), ptr(d_descR), ptr(d_monoR), ptr(d_nR), B, cap,
                                            ptr(idx2), ptr(dist2), ptr(match), st))
        return idx2, dist2, match
