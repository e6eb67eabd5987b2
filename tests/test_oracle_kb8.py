"""CPU: the KB8 restatement (oracle/kb8_oracle.cpp) against the reference's own KannalaBrandt8::project (both float
overloads) and ::unproject bodies, compiled verbatim into oracle/_ref/libref_kb8.so (oracle/ref_build.sh) -- bit-exact.
TriangulateMatches has no compiled reference (Eigen::JacobiSVD; Eigen is not in the image): its restatement is checked
through properties the reference's code implies (known 3-D points come back, each rejection code is reachable)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "oracle", "_ref", "libref_kb8.so")
# TUM-VI style fisheye intrinsics (Examples/Stereo-Inertial/TUM-VI.yaml shape: fx fy cx cy k1..k4) and a second camera
P1 = np.array([190.978477, 190.973307, 254.931706, 256.897442, 0.003482389, 0.000715034, -0.002053236, 0.000202936], np.float32)
P2 = np.array([190.442369, 190.434448, 252.598029, 254.917267, 0.003400724, 0.001766232, -0.002663594, 0.000329930], np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libref_kb8.so not built (needs /root/reference at build time)")
    return C.CDLL(REF)


def _points(rng, n):
    p = rng.normal(size=(n, 3)).astype(np.float32) * np.float32(3)
    p[: n // 2, 2] = np.abs(p[: n // 2, 2]) + np.float32(0.2)      # half of them in front of the camera
    p[0] = (0, 0, 1)                                                # on the axis
    p[1] = (1, 0, 0)                                                # 90 degrees off the axis
    return np.ascontiguousarray(p)


@pytest.mark.parametrize("seed", range(4))
def test_project_matches_the_reference_body(ref, seed):
    rng = np.random.default_rng(seed)
    p3 = _points(rng, 4000)
    for P in (P1, P2):
        a = O.kb8_project(P, p3)
        for fn in (ref.ref_kb8_project, ref.ref_kb8_project_eig):
            b = np.empty_like(a)
            fn(_p(P), _p(p3), len(p3), _p(b))
            assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("seed", range(4))
def test_unproject_matches_the_reference_body(ref, seed):
    rng = np.random.default_rng(100 + seed)
    uv = (rng.random((4000, 2)) * 512).astype(np.float32)
    uv[0] = P1[2:4]                                                 # the principal point: theta_d == 0 branch
    uv[1] = (5000, -3000)                                           # far outside: theta_d clamped to pi/2
    for P in (P1, P2):
        for prec in (1e-6, 1e-3):
            a = O.kb8_unproject(P, uv, prec)
            b = np.empty_like(a)
            ref.ref_kb8_unproject(_p(P), C.c_float(prec), _p(uv), len(uv), _p(b))
            assert a.tobytes() == b.tobytes()


def test_unproject_inverts_project():
    rng = np.random.default_rng(7)
    p3 = _points(rng, 2000)[:1000]                                  # in front of the camera
    p3 = p3[np.hypot(p3[:, 0], p3[:, 1]) < 5 * p3[:, 2]]            # within ~79 degrees of the axis
    uv = O.kb8_project(P1, p3)
    ray = O.kb8_unproject(P1, uv)
    assert np.allclose(ray[:, :2], p3[:, :2] / p3[:, 2:3], rtol=2e-3, atol=2e-4)


def _rig(rng, n):
    """Points in front of a 10 cm fisheye stereo rig; R12 / t12 as Frame::mRlr / mtlr (camera 2 -> camera 1)."""
    ang = 0.02
    R12 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    t12 = np.array([0.1, 0.002, -0.001], np.float32)
    X1 = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(0.5, 3, n)], 1).astype(np.float32)
    X2 = (X1 - t12) @ R12                                           # R21 (X1 - t12)
    return R12, t12, X1, np.ascontiguousarray(X2.astype(np.float32))


def test_triangulate_recovers_known_points_and_rejects():
    rng = np.random.default_rng(3)
    R12, t12, X1, X2 = _rig(rng, 500)
    pt1, pt2 = O.kb8_project(P1, X1), O.kb8_project(P2, X2)
    s = np.ones(len(X1), np.float32)
    depth, p3d = O.kb8_triangulate(P1, P2, R12, t12, pt1, pt2, s, s)
    ok = depth > 0
    assert ok.mean() > 0.95
    assert np.allclose(p3d[ok], X1[ok], rtol=2e-2, atol=2e-2) and np.allclose(depth[ok], X1[ok, 2], rtol=2e-2, atol=2e-2)
    far = np.tile(np.array([[0.3, 0.2, 4000.0]], np.float32), (4, 1))           # no parallax
    d, _ = O.kb8_triangulate(P1, P2, R12, t12, O.kb8_project(P1, far), O.kb8_project(P2, (far - t12) @ R12), s[:4], s[:4])
    assert (d == -1).all()
    bad = pt2.copy()
    bad[:, 1] += 25                                                 # off the epipolar curve: reprojection error
    d, _ = O.kb8_triangulate(P1, P2, R12, t12, pt1, bad, s, s)
    assert set(np.unique(d[d < 0]).tolist()) <= {-1.0, -2.0, -3.0, -4.0, -5.0} and (d < 0).mean() > 0.9
    assert (d == -4).any() or (d == -5).any()


def test_oracle_matches_the_committed_reference_outputs():
    """tests/golden/kb8_ref.npz holds outputs of the reference's own project / unproject (make_golden_kb8.py): this
    check needs neither /root/reference nor oracle/_ref."""
    g = np.load(os.path.join(HERE, "golden", "kb8_ref.npz"))
    for name in ("1", "2"):
        P = g["P" + name]
        assert O.kb8_project(P, g["p3d"]).tobytes() == g["project" + name].tobytes()
        assert O.kb8_unproject(P, g["uv"]).tobytes() == g["unproject" + name].tobytes()


def test_null_vector_step_against_lapack():
    """The restated replacement of Eigen::JacobiSVD<Matrix4f> (KannalaBrandt8.cpp:566-568) against LAPACK's SVD: the
    unit vector it returns is the right singular vector of the smallest singular value (|A x| = sigma_min), and where
    that value is well separated it is LAPACK's vector up to sign."""
    rng = np.random.default_rng(11)
    n = 4000
    A = rng.normal(size=(n, 4, 4)).astype(np.float32)
    A[-1, 3] = A[-1, 2]                                                      # rank deficient
    x = O.kb8_null_vectors(A)
    A64 = A.astype(np.float64)
    _, sv, vt = np.linalg.svd(A64)
    res = np.linalg.norm(np.einsum("nij,nj->ni", A64, x), axis=1)
    assert np.allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-12)
    assert np.allclose(res, sv[:, 3], rtol=1e-6, atol=1e-9)
    sep = sv[:, 2] - sv[:, 3] > 1e-3
    dot = np.abs(np.einsum("nj,nj->n", x, vt[:, 3]))
    assert sep.mean() > 0.95 and np.allclose(dot[sep], 1.0, atol=1e-8)
