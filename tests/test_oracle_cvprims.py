"""Pin the oracle's restated OpenCV primitives (oracle/cvprims.cpp) against cv2 4.13.0 --
the OpenCV the reference links but does not vendor.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)


def _img(h, w, seed, smooth=True):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w)).astype(np.uint8)
    if smooth:
        a = cv2.GaussianBlur(a, (5, 5), 1.2)
        a = cv2.normalize(a, None, 0, 255, cv2.NORM_MINMAX)
    return a


@pytest.mark.parametrize("sw,sh,dw,dh", [(752, 480, 627, 400), (627, 400, 522, 333), (512, 512, 427, 427),
                                          (1280, 720, 1067, 600), (252, 161, 210, 134), (100, 60, 50, 30),
                                          (37, 29, 31, 24)])
def test_resize_linear(sw, sh, dw, dh):
    src = _img(sh, sw, sw + dh)
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(O.resize_linear(src, dw, dh), ref)


@pytest.mark.parametrize("w,h", [(64, 48), (21, 20), (210, 134)])
def test_border101(w, h):
    src = _img(h, w, 3)
    ref = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    assert np.array_equal(O.border101(src, 19), ref)


@pytest.mark.parametrize("th", [7, 20])
@pytest.mark.parametrize("seed,smooth", [(0, True), (1, True), (2, False)])
def test_fast(th, seed, smooth):
    img = _img(61, 47, seed, smooth)
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(img)
    ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
    got = O.fast(img, th, True)
    assert np.array_equal(got, ref)


def test_fast_no_nms_and_tiny():
    img = _img(40, 40, 5)
    det = cv2.FastFeatureDetector_create(threshold=12, nonmaxSuppression=False,
                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = np.array([[int(k.pt[0]), int(k.pt[1])] for k in det.detect(img)], np.int32).reshape(-1, 2)
    assert np.array_equal(O.fast(img, 12, False)[:, :2], ref)
    assert len(O.fast(_img(6, 30, 1), 7)) == 0


@pytest.mark.parametrize("w,h", [(752, 480), (210, 134), (33, 9)])
def test_blur7(w, h):
    src = _img(h, w, 11, smooth=False)
    ref = cv2.GaussianBlur(src, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(O.blur7(src), ref)


def test_blur7_impulse_known_answer():
    src = np.zeros((15, 15), np.uint8)
    src[7, 7] = 255
    out = O.blur7(src).astype(np.int64)
    k = np.array([18, 34, 48, 56, 48, 34, 18])
    exp = (np.outer(k, k) * 255 + 32768) >> 16
    assert np.array_equal(out[4:11, 4:11], exp)


def test_fast_atan2():
    rng = np.random.default_rng(0)
    y = rng.integers(-40000, 40000, 200000).astype(np.float32)
    x = rng.integers(-40000, 40000, 200000).astype(np.float32)
    y[:4] = [1, 3, 0, 0]
    x[:4] = [1, -5, 0, -1]
    # the scalar cv::fastAtan2(float, float) is what the reference calls (ORBextractor.cc:137);
    # cv2.cartToPolar's SIMD path contracts to FMA and differs in the last ulp, so it is not used.
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    got = O.fast_atan2(y, x)
    assert np.array_equal(got, ref)
    assert got[2] == 0.0 and got[3] == 180.0


def test_bf_knn2_and_ties():
    rng = np.random.default_rng(4)
    q = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    t = rng.integers(0, 256, (500, 32)).astype(np.uint8)
    t[77] = t[13]          # duplicate rows: tie -> lower train index first
    t[400] = q[5]
    t[20] = q[5]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    ref = bf.knnMatch(q, t, k=2)
    idx, dist = O.knn2(q, t)
    for i, pair in enumerate(ref):
        assert [m.trainIdx for m in pair] == list(idx[i])
        assert [int(m.distance) for m in pair] == list(dist[i])
    assert list(idx[5]) == [20, 400]


def test_hamming_matches_popcount():
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (64, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (64, 32)).astype(np.uint8)
    for i in range(64):
        assert O.hamming(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())
    assert O.hamming(a[0], a[0]) == 0
    assert O.hamming(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


@pytest.mark.parametrize("code,channels,rgb", [("COLOR_BGR2GRAY", 3, False), ("COLOR_RGB2GRAY", 3, True),
                                               ("COLOR_BGRA2GRAY", 4, False), ("COLOR_RGBA2GRAY", 4, True)])
def test_cvt_gray(code, channels, rgb):
    """cv::cvtColor to gray as Tracking::GrabImage* calls it (src/Tracking.cc:1563-1590)."""
    rng = np.random.default_rng(channels + rgb)
    img = rng.integers(0, 256, (61, 83, channels)).astype(np.uint8)
    img[0, :8] = [[0] * channels, [255] * channels] * 4
    assert np.array_equal(O.cvt_gray(img, rgb), cv2.cvtColor(img, getattr(cv2, code)))


def _rectify_like_maps(dh, dw, sh, sw, seed):
    """Maps shaped like initUndistortRectifyMap output: smooth radial distortion plus a small rotation, reaching
    outside the source near the corners."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:dh, 0:dw].astype(np.float64)
    xn, yn = (x - dw / 2) / (0.9 * dw), (y - dh / 2) / (0.9 * dw)
    r2 = xn * xn + yn * yn
    k1, k2 = rng.uniform(-0.4, 0.4), rng.uniform(-0.1, 0.2)
    th = rng.uniform(-0.03, 0.03)
    f = 1 + k1 * r2 + k2 * r2 * r2
    xd, yd = xn * f, yn * f
    mx = (np.cos(th) * xd - np.sin(th) * yd) * 0.9 * dw * sw / dw + sw / 2
    my = (np.sin(th) * xd + np.cos(th) * yd) * 0.9 * dw * sh / dh + sh / 2
    return mx.astype(np.float32), my.astype(np.float32)


@pytest.mark.parametrize("dh,dw,sh,sw,seed", [(480, 752, 480, 752, 0), (100, 131, 90, 140, 1), (64, 64, 200, 300, 2)])
def test_remap_linear(dh, dw, sh, sw, seed):
    """cv::remap(.., INTER_LINEAR) with CV_32FC1 maps and the default constant border (src/System.cc:292-293)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
    mx, my = _rectify_like_maps(dh, dw, sh, sw, seed)
    assert np.array_equal(O.remap_linear(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR))
    # exact half positions (round-half-even of x*32), integer positions, far outside, negative fractions
    mx2 = (rng.integers(-40 * 64, (sw + 40) * 64, (dh, dw)) / 64.0).astype(np.float32)
    my2 = (rng.integers(-40 * 64, (sh + 40) * 64, (dh, dw)) / 64.0).astype(np.float32)
    mx2[0, :4] = [-1e7, 1e7, -0.5, sw - 0.5]
    assert np.array_equal(O.remap_linear(src, mx2, my2), cv2.remap(src, mx2, my2, cv2.INTER_LINEAR))


@pytest.mark.parametrize("dist", [[-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05],          # EuRoC cam0
                                  [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0021],  # with k3
                                  [0.262383, -0.953104, -0.005358, 0.002628, 1.163314],           # TUM1
                                  [-3.5, 0.2, 0.01, -0.02]])                                      # icdist < 0 near the borders
def test_undistort_points(dist):
    """cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) as Frame::UndistortKeyPoints calls it
    (src/Frame.cc:1025).  The camera matrix is float (cv::Mat CV_32F), the arithmetic double."""
    rng = np.random.default_rng(len(dist))
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float32)
    D = np.array(dist, np.float32)
    pts = np.stack([rng.uniform(0, 752, 3000), rng.uniform(0, 480, 3000)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [751, 0], [0, 479], [751, 479]]
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    got = O.undistort_points(pts, (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), D)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
