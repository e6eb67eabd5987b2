"""The OpenCV calls ORB-SLAM3 makes on a frame before ORBextractor, on the device (csrc/intake.cu):
cv::cvtColor to gray (Tracking.cc:1563-1590), cv::remap stereo rectification (System.cc:286-293) and cv::resize
(System.cc:295-297).  Bit-exact with cv2 4.13 for 8-bit images."""
import numpy as np

from ._lib import check, lib, ptr


def cvtColorToGray(img, rgb=False, device=0):
    """cv::cvtColor(img, gray, COLOR_{BGR,RGB,BGRA,RGBA}2GRAY); rgb = the first channel is red (Tracking::mbRGB)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, c = img.shape
    out = np.empty((h, w), np.uint8)
    check(lib().orbfe_cvt_gray(ptr(img), h, w, img.strides[0], c, int(rgb), ptr(out), w, device))
    return out


def remap(src, map_x, map_y, device=0):
    """cv::remap(src, dst, M1, M2, cv::INTER_LINEAR) with CV_32FC1 maps and the default constant (0) border."""
    src = np.ascontiguousarray(src, np.uint8)
    map_x = np.ascontiguousarray(map_x, np.float32)
    map_y = np.ascontiguousarray(map_y, np.float32)
    dh, dw = map_x.shape
    out = np.empty((dh, dw), np.uint8)
    check(lib().orbfe_remap_linear(ptr(src), src.shape[0], src.shape[1], src.strides[0], ptr(map_x), ptr(map_y), dh, dw,
                                   ptr(out), dw, device))
    return out


def resize(src, dsize, device=0):
    """cv::resize(src, dst, dsize) (INTER_LINEAR); dsize = (width, height) as in OpenCV."""
    src = np.ascontiguousarray(src, np.uint8)
    dw, dh = dsize
    out = np.empty((dh, dw), np.uint8)
    check(lib().orbfe_resize_linear(ptr(src), src.shape[0], src.shape[1], src.strides[0], dh, dw, ptr(out), dw, device))
    return out


def undistortKeyPoints(keys, K, dist_coef, device=0):
    """Frame::UndistortKeyPoints (Frame.cc:1003-1051): mvKeys -> mvKeysUn.  K = (fx, fy, cx, cy)."""
    from ._lib import KP_DTYPE
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    d = np.ascontiguousarray(dist_coef, np.float32)
    out = np.empty_like(keys)
    check(lib().orbfe_undistort_keypoints(ptr(keys), len(keys), float(K[0]), float(K[1]), float(K[2]), float(K[3]), ptr(d),
                                          len(d), ptr(out), device))
    return out


# ---- device-resident forms: torch CUDA tensors in and out, enqueued on a stream, no synchronisation -------------
def cvtColorToGray_device(d_img, d_out, rgb=False, stream=None):
    """d_img: (rows, cols, 3|4) uint8 CUDA tensor, d_out: (rows, cols) uint8 CUDA tensor."""
    import ctypes as C
    rows, cols, ch = d_img.shape
    st = C.c_void_p(stream.cuda_stream) if stream is not None else None
    check(lib().orbfe_cvt_gray_device(ptr(d_img), rows, cols, d_img.stride(0), ch, int(rgb), ptr(d_out), d_out.stride(0), st))


def remap_device(d_src, d_map_x, d_map_y, d_out, stream=None):
    """d_src: (rows, cols) uint8, d_map_x / d_map_y: (drows, dcols) float32, d_out: (drows, dcols) uint8; CUDA tensors."""
    import ctypes as C
    st = C.c_void_p(stream.cuda_stream) if stream is not None else None
    drows, dcols = d_map_x.shape
    check(lib().orbfe_remap_linear_device(ptr(d_src), d_src.shape[0], d_src.shape[1], d_src.stride(0), ptr(d_map_x),
                                          ptr(d_map_y), d_map_x.stride(0), drows, dcols, ptr(d_out), d_out.stride(0), st))
