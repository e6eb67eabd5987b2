"""ORB_SLAM3::KannalaBrandt8 (/root/reference/include/CameraModels/KannalaBrandt8.h, src/CameraModels/KannalaBrandt8.cpp):
the fisheye camera geometry Frame::ComputeStereoFishEyeMatches applies to every kNN match (src/Frame.cc:1560-1587),
batched on the device (csrc/kb8.cu).  Same method names and argument meaning as the reference; arrays instead of single
points / keypoints."""
import numpy as np

from ._lib import check, lib, ptr


class KannalaBrandt8:
    def __init__(self, vParameters, precision=1e-6, device=0):
        """vParameters = {fx, fy, cx, cy, k1, k2, k3, k4} (KannalaBrandt8.h:47-57)."""
        self.mvParameters = np.ascontiguousarray(vParameters, np.float32)
        if self.mvParameters.shape != (8,):
            raise ValueError("KannalaBrandt8 takes 8 parameters")
        self.precision = float(precision)
        self.device = device

    def GetPrecision(self):
        return self.precision

    def project(self, p3D):
        """n x 3 points in the camera frame -> n x 2 pixels (.cpp:40-55, 84-101)."""
        p3D = np.ascontiguousarray(p3D, np.float32).reshape(-1, 3)
        uv = np.empty((len(p3D), 2), np.float32)
        check(lib().orbfe_kb8_project(ptr(self.mvParameters), ptr(p3D), len(p3D), ptr(uv), self.device))
        return uv

    def unproject(self, p2D):
        """n x 2 pixels -> n x 3 rays (x, y, 1) (.cpp:180-217)."""
        p2D = np.ascontiguousarray(p2D, np.float32).reshape(-1, 2)
        rays = np.empty((len(p2D), 3), np.float32)
        check(lib().orbfe_kb8_unproject(ptr(self.mvParameters), self.precision, ptr(p2D), len(p2D), ptr(rays), self.device))
        return rays

    def TriangulateMatches(self, pCamera2, pt1, pt2, R12, t12, sigmaLevel, unc):
        """Per match i: (depth[i], p3D[i]) as KannalaBrandt8::TriangulateMatches(pCamera2, kp1, kp2, R12, t12,
        sigmaLevel, unc, p3D) returns / writes them (.cpp:439-515); pt1 / pt2 = kp.pt of the two keypoints,
        sigmaLevel / unc = scalars or per-match arrays.  p3D rows of rejected matches are NaN."""
        pt1 = np.ascontiguousarray(pt1, np.float32).reshape(-1, 2)
        pt2 = np.ascontiguousarray(pt2, np.float32).reshape(-1, 2)
        n = len(pt1)
        if len(pt2) != n:
            raise ValueError("pt1 and pt2 differ in length")
        s1 = np.ascontiguousarray(np.broadcast_to(np.asarray(sigmaLevel, np.float32), (n,)))
        s2 = np.ascontiguousarray(np.broadcast_to(np.asarray(unc, np.float32), (n,)))
        R = np.ascontiguousarray(R12, np.float32).reshape(3, 3)
        t = np.ascontiguousarray(t12, np.float32).reshape(3)
        depth = np.empty(n, np.float32)
        p3d = np.empty((n, 3), np.float32)
        check(lib().orbfe_kb8_triangulate_matches(ptr(self.mvParameters), self.precision, ptr(pCamera2.mvParameters),
                                                  pCamera2.precision, ptr(R), ptr(t), ptr(pt1), ptr(pt2), ptr(s1), ptr(s2),
                                                  n, ptr(depth), ptr(p3d), self.device))
        return depth, p3d

    def epipolarConstrain(self, pCamera2, pt1, pt2, R12, t12, sigmaLevel, unc):
        """.cpp:322-328: TriangulateMatches(...) > 0.0001f, per match."""
        return self.TriangulateMatches(pCamera2, pt1, pt2, R12, t12, sigmaLevel, unc)[0] > np.float32(0.0001)
