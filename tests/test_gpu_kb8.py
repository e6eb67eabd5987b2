"""GPU: orbfe_kb8_project / _unproject / _triangulate_matches (csrc/kb8.cu) against the oracle's restatement of
KannalaBrandt8::project / unproject / TriangulateMatches (oracle/kb8_oracle.cpp; project and unproject pinned against
the reference's own bodies, tests/test_oracle_kb8.py).  Floating point: the tolerances below are the bar.
  project    |du|, |dv| < 1e-3 px   (transcendentals evaluated in double on the device, glibc's float ones in the oracle)
  unproject  <= 8 ulp on every ray component
  triangulate: accept / reject equal on > 99.8 % of the matches (a threshold can flip on a 1-ulp ray), identical
               rejection codes where both reject, depth and p3D within 2e-4 relative where both accept."""
import numpy as np
import pytest

from oracle import oracle as O
from test_oracle_kb8 import P1, P2, _points, _rig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orbfe():
    import orbfe as m
    m.lib()
    return m


def _ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def test_project_unproject(orbfe):
    rng = np.random.default_rng(0)
    p3 = _points(rng, 50000)
    uv = (rng.random((50000, 2)) * 512).astype(np.float32)
    uv[0] = P1[2:4]                 # principal point: theta_d == 0 branch, ray (0, 0, 1)
    uv[1] = (5000, -3000)           # theta_d clamped to pi/2
    for P in (P1, P2):
        cam = orbfe.KannalaBrandt8(P)
        a, b = O.kb8_project(P, p3), cam.project(p3)
        assert np.abs(a - b).max() < 1e-3 and (a == b).mean() > 0.8
        for prec in (1e-6, 1e-3):
            cam = orbfe.KannalaBrandt8(P, prec)
            a, b = O.kb8_unproject(P, uv, prec), cam.unproject(uv)
            assert _ulps(a, b).max() <= 8 and (a == b).mean() > 0.95
            assert b[0].tolist() == [0.0, 0.0, 1.0] or P is P2
    assert orbfe.KannalaBrandt8(P1).project(np.zeros((0, 3), np.float32)).shape == (0, 2)
    with pytest.raises(ValueError):
        orbfe.KannalaBrandt8(P1[:5])


def test_triangulate_matches(orbfe):
    rng = np.random.default_rng(5)
    R12, t12, X1, X2 = _rig(rng, 20000)
    pt1, pt2 = O.kb8_project(P1, X1), O.kb8_project(P2, X2)
    pt2[::3] += rng.normal(scale=1.5, size=pt2[::3].shape).astype(np.float32)
    pt2[5::50] += 40
    sig = rng.choice(np.float32([1.0, 1.44, 2.0736, 2.985984]), len(X1))
    unc = rng.choice(np.float32([1.0, 1.44, 2.0736, 2.985984]), len(X1))
    d_o, p_o = O.kb8_triangulate(P1, P2, R12, t12, pt1, pt2, sig, unc)
    c1, c2 = orbfe.KannalaBrandt8(P1), orbfe.KannalaBrandt8(P2)
    d_g, p_g = c1.TriangulateMatches(c2, pt1, pt2, R12, t12, sig, unc)
    assert len(set(np.unique(d_o[d_o < 0]).tolist())) >= 3
    same = (d_o < 0) == (d_g < 0)
    assert same.mean() > 0.998
    neg = same & (d_o < 0)
    assert (d_o[neg] == d_g[neg]).mean() > 0.998
    pos = same & (d_o > 0)
    assert pos.sum() > 5000
    assert np.allclose(d_o[pos], d_g[pos], rtol=2e-4, atol=1e-5) and np.allclose(p_o[pos], p_g[pos], rtol=2e-4, atol=2e-5)
    assert np.isnan(p_g[d_g < 0]).all()
    # the known points come back, and epipolarConstrain is the thresholded form
    clean = pos & (np.arange(len(X1)) % 3 != 0)
    assert np.allclose(p_g[clean], X1[clean], rtol=2e-2, atol=2e-2)
    ok = c1.epipolarConstrain(c2, pt1, pt2, R12, t12, sig, unc)
    assert np.array_equal(ok, d_g > np.float32(0.0001))
    # scalar sigma / unc broadcast
    d_s, _ = c1.TriangulateMatches(c2, pt1[:100], pt2[:100], R12, t12, 1.0, 1.0)
    d_v, _ = c1.TriangulateMatches(c2, pt1[:100], pt2[:100], R12, t12, np.ones(100, np.float32), np.ones(100, np.float32))
    assert np.array_equal(d_s, d_v)


def test_against_the_reference_outputs(orbfe):
    """The CUDA path against outputs of the reference's own KannalaBrandt8::project / unproject
    (tests/golden/kb8_ref.npz), no oracle in between; same tolerances as above."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kb8_ref.npz"))
    for name in ("1", "2"):
        cam = orbfe.KannalaBrandt8(g["P" + name])
        assert np.abs(cam.project(g["p3d"]) - g["project" + name]).max() < 1e-3
        assert _ulps(cam.unproject(g["uv"]), np.ascontiguousarray(g["unproject" + name])).max() <= 8


def test_null_vector_step_is_bit_exact(orbfe):
    """The step that replaces Eigen::JacobiSVD<Matrix4f> (KannalaBrandt8.cpp:566-568; Eigen is not in the image, so this
    step is "parity unpinned" against the reference): the CUDA kernel and the oracle restate the same fp64 cyclic Jacobi
    and must agree to the last bit -- triangulation-shaped systems, random ones, rank-deficient and zero matrices."""
    import ctypes as C
    from orbfe import _lib
    rng = np.random.default_rng(7)
    n = 20000
    A = rng.normal(size=(n, 4, 4)).astype(np.float32)
    r = rng.normal(size=(n // 2, 2, 2)).astype(np.float32)                  # rows x * T.row(2) - T.row(0): the reference's shape
    T2 = np.concatenate([np.tile(np.eye(3, dtype=np.float32), (n // 2, 1, 1)), rng.normal(size=(n // 2, 3, 1)).astype(np.float32)], 2)
    T1 = np.tile(np.concatenate([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)], 1), (n // 2, 1, 1))
    for k, (T, j) in enumerate(((T1, 0), (T1, 0), (T2, 1), (T2, 1))):
        A[: n // 2, k] = r[:, j, k % 2, None] * T[:, 2] - T[:, k % 2]
    A[-1] = 0
    A[-2, 3] = A[-2, 2]                                                        # rank deficient
    A[-3] = np.eye(4)
    x = np.empty((n, 4), np.float64)
    _lib.check(_lib.lib().orbfe_debug_kb8_null_vectors(_lib.ptr(np.ascontiguousarray(A)), n, _lib.ptr(x), 0))
    ex = O.kb8_null_vectors(A)
    assert np.array_equal(x.view(np.uint64), ex.view(np.uint64))
    # and it is the direction of the smallest singular value: |A x| = sigma_min for a unit x
    sel = slice(n // 2, n // 2 + 2000)
    res = np.linalg.norm(np.einsum("nij,nj->ni", A[sel].astype(np.float64), x[sel]), axis=1)
    sv = np.linalg.svd(A[sel].astype(np.float64), compute_uv=False)
    assert np.allclose(np.linalg.norm(x[sel], axis=1), 1.0, atol=1e-12) and np.allclose(res, sv[:, 3], rtol=1e-6, atol=1e-9)
