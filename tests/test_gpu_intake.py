"""GPU parity of the frame-intake kernels (csrc/intake.cu: cvtColor to gray, remap, resize) against the oracle's
restatement AND against cv2 4.13 itself (the OpenCV the reference links)."""
import os
import sys

import numpy as np
import pytest

import synth
from oracle import oracle as O
from test_oracle_cvprims import _rectify_like_maps

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def orbfe():
    sys.path.insert(0, os.path.join(ROOT, "orb-slam3_byzyh_b200"))
    import orbfe as m
    return m


def _cv2():
    try:
        import cv2
        return cv2
    except ImportError:
        return None


@pytest.mark.parametrize("channels,rgb,code", [(3, False, "COLOR_BGR2GRAY"), (3, True, "COLOR_RGB2GRAY"),
                                               (4, False, "COLOR_BGRA2GRAY"), (4, True, "COLOR_RGBA2GRAY")])
def test_cvt_gray(orbfe, channels, rgb, code):
    rng = np.random.default_rng(channels * 2 + rgb)
    img = rng.integers(0, 256, (480, 752, channels)).astype(np.uint8)
    got = orbfe.intake.cvtColorToGray(img, rgb)
    assert np.array_equal(got, O.cvt_gray(img, rgb))
    cv2 = _cv2()
    if cv2 is not None:
        assert np.array_equal(got, cv2.cvtColor(img, getattr(cv2, code)))
    small = img[:7, :5].copy()                               # odd sizes, strided source
    assert np.array_equal(orbfe.intake.cvtColorToGray(small, rgb), O.cvt_gray(small, rgb))


@pytest.mark.parametrize("dh,dw,sh,sw,seed", [(480, 752, 480, 752, 0), (100, 131, 90, 140, 1), (720, 1280, 800, 1400, 2)])
def test_remap(orbfe, dh, dw, sh, sw, seed):
    rng = np.random.default_rng(seed)
    src = synth.synth_frame(sh, sw, seed)
    mx, my = _rectify_like_maps(dh, dw, sh, sw, seed)
    got = orbfe.intake.remap(src, mx, my)
    assert np.array_equal(got, O.remap_linear(src, mx, my))
    mx2 = (rng.integers(-40 * 64, (sw + 40) * 64, (dh, dw)) / 64.0).astype(np.float32)   # halves, integers, outside
    my2 = (rng.integers(-40 * 64, (sh + 40) * 64, (dh, dw)) / 64.0).astype(np.float32)
    mx2[0, :6] = [-1e7, 1e7, -0.5, sw - 0.5, np.nan, np.inf]
    got2 = orbfe.intake.remap(src, mx2, my2)
    assert np.array_equal(got2, O.remap_linear(src, mx2, my2))
    cv2 = _cv2()
    if cv2 is not None:
        assert np.array_equal(got, cv2.remap(src, mx, my, cv2.INTER_LINEAR))
        ok = np.isfinite(mx2)
        assert np.array_equal(got2[ok], cv2.remap(src, mx2, my2, cv2.INTER_LINEAR)[ok])


@pytest.mark.parametrize("sw,sh,dw,dh", [(752, 480, 600, 350), (1280, 720, 640, 360), (640, 480, 752, 480), (100, 60, 300, 180),
                                          (333, 222, 333, 222)])
def test_resize(orbfe, sw, sh, dw, dh):
    src = synth.synth_frame(sh, sw, sw + dh)
    got = orbfe.intake.resize(src, (dw, dh))
    assert np.array_equal(got, O.resize_linear(src, dw, dh))
    cv2 = _cv2()
    if cv2 is not None:
        assert np.array_equal(got, cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_rectify_then_extract_equals_reference_chain(orbfe):
    """The stereo intake chain of System::TrackStereo: remap on the device, then ORBextractor, equals
    cv-style remap (oracle) followed by the oracle extractor."""
    src = synth.synth_frame(480, 752, 5)
    mx, my = _rectify_like_maps(480, 752, 480, 752, 3)
    rect = orbfe.intake.remap(src, mx, my)
    ex = orbfe.ORBextractor(1000)
    _, k, d = ex(rect, None, (0, 0))
    oex = O.Extractor(1000)
    _, ok, od = oex(O.remap_linear(src, mx, my), (0, 0))
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od) and len(k) > 500


@pytest.mark.parametrize("dist", [[-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05],
                                  [0.262383, -0.953104, -0.005358, 0.002628, 1.163314],
                                  [-3.5, 0.2, 0.01, -0.02], [0.0, 0.1, 0.0, 0.0]])
def test_undistort_keypoints(orbfe, dist):
    """Frame::UndistortKeyPoints: coordinates bit-equal to cv::undistortPoints, other fields untouched; a zero first
    coefficient copies (Frame.cc:1005-1009)."""
    from oracle.oracle import KP_DTYPE
    rng = np.random.default_rng(7)
    n = 3000
    keys = np.zeros(n, KP_DTYPE)
    keys["x"], keys["y"] = rng.uniform(0, 752, n), rng.uniform(0, 480, n)
    keys["size"], keys["angle"], keys["response"] = 31.0, rng.uniform(0, 360, n), rng.uniform(1, 200, n)
    keys["octave"], keys["class_id"] = rng.integers(0, 8, n), -1
    K = tuple(np.float32(v) for v in (458.654, 457.296, 367.215, 248.375))
    got = orbfe.intake.undistortKeyPoints(keys, K, dist)
    exp = keys.copy()
    if dist[0] != 0.0:
        xy = O.undistort_points(np.stack([keys["x"], keys["y"]], 1), K, dist)
        exp["x"], exp["y"] = xy[:, 0], xy[:, 1]
        cv2 = _cv2()
        if cv2 is not None:
            Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
            ref = cv2.undistortPoints(np.stack([keys["x"], keys["y"]], 1).reshape(-1, 1, 2), Km, np.array(dist, np.float32), None,
                                      Km).reshape(-1, 2)
            assert np.array_equal(got["x"].view(np.uint32), ref[:, 0].view(np.uint32))
            assert np.array_equal(got["y"].view(np.uint32), ref[:, 1].view(np.uint32))
    assert got.tobytes() == exp.tobytes()


def test_device_resident_chain(orbfe):
    """A colour stereo frame uploaded once: cvtColor -> remap -> ORBextractor, all with device pointers on one stream
    (no host round trip), equals the host chain of the oracle."""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    gray = synth.synth_frame(480, 752, 9)
    bgr = np.stack([np.clip(gray.astype(np.int32) + rng.integers(-20, 21, gray.shape), 0, 255).astype(np.uint8) for _ in range(3)], 2)
    mx, my = _rectify_like_maps(480, 752, 480, 752, 6)
    d_bgr = torch.from_numpy(bgr).to(dev)
    d_mx, d_my = torch.from_numpy(mx).to(dev), torch.from_numpy(my).to(dev)
    d_gray = torch.empty((480, 752), dtype=torch.uint8, device=dev)
    d_rect = torch.empty((1, 480, 752), dtype=torch.uint8, device=dev)
    ex = orbfe.ORBextractor(1000)
    cap = 1000 + 64 * 8 + 64
    d_kps = torch.empty((1, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((1, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.empty(1, dtype=torch.int32, device=dev)
    d_mono = torch.empty(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    orbfe.intake.cvtColorToGray_device(d_bgr, d_gray, rgb=False, stream=st)
    orbfe.intake.remap_device(d_gray, d_mx, d_my, d_rect[0], stream=st)
    ex.extract_batch_device(d_rect, (0, 0), d_kps, d_desc, d_n, d_mono, stream=st)
    torch.cuda.synchronize()
    n = int(d_n[0])
    from oracle.oracle import KP_DTYPE
    kps = d_kps[0, :n].cpu().numpy().reshape(-1).view(KP_DTYPE)
    desc = d_desc[0, :n].cpu().numpy()
    orect = O.remap_linear(O.cvt_gray(bgr, False), mx, my)
    assert np.array_equal(d_rect[0].cpu().numpy(), orect)
    _, ok, od = O.Extractor(1000)(orect, (0, 0))
    assert n == len(ok) and n > 500 and kps.tobytes() == ok.tobytes() and np.array_equal(desc, od)


def test_rectification_fused_into_the_extractor(orbfe):
    """orbfe_extractor_set_rectification: raw frames in, pyramid level 0 = cv::remap(frame) + border, everything else as
    ORBextractor on the rectified image (System::TrackStereo, System.cc:286-293); single frames, batches, raw size
    different from the rectified size, and switching it off again."""
    raw_h, raw_w, h, w = 520, 800, 480, 752
    frames = np.stack([synth.synth_frame(raw_h, raw_w, 80 + i) for i in range(5)])
    mx, my = _rectify_like_maps(h, w, raw_h, raw_w, 8)
    ex = orbfe.ORBextractor(1000)
    oex = O.Extractor(1000)
    ex.set_rectification(mx, my)
    rect0 = O.remap_linear(frames[0], mx, my)
    mono, k, d = ex(frames[0], None, (0, 1000))
    omono, ok, od = oex(rect0, (0, 1000))
    assert mono == omono and k.tobytes() == ok.tobytes() and np.array_equal(d, od) and len(k) > 500
    assert np.array_equal(ex.pyramid_level(0), rect0)                       # mvImagePyramid[0] is the rectified image
    assert np.array_equal(ex.pyramid_level(1, with_border=True), oex.level(1)["padded"])
    n, monos, kps, desc = ex.extract_batch(frames, (0, 1000))
    for i in range(len(frames)):
        omono, ok, od = oex(O.remap_linear(frames[i], mx, my), (0, 1000))
        assert monos[i] == omono and n[i] == len(ok) and kps[i, :n[i]].tobytes() == ok.tobytes()
        assert np.array_equal(desc[i, :n[i]], od)
    ex.set_rectification(None)
    mono, k, d = ex(frames[0], None, (0, 1000))
    omono, ok, od = oex(frames[0], (0, 1000))
    assert mono == omono and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
