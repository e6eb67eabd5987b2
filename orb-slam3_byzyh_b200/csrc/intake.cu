// intake.cu -- the image operations ORB-SLAM3 runs on a frame BEFORE ORBextractor, so that a
// camera frame can go H2D once and stay in HBM through rectification, extraction and matching:
//   cv::cvtColor(.., COLOR_{RGB,BGR,RGBA,BGRA}2GRAY)   /root/reference/src/Tracking.cc:1563-1590, 1623-1636, 1702-1716
//   cv::remap(im, out, M1, M2, cv::INTER_LINEAR)       src/System.cc:286-293 (stereo rectification, Settings::needToRectify)
//   cv::resize(im, out, newImSize)                     src/System.cc:295-297, 371-376, 457-459 (Settings::needToResize)
// OpenCV is not vendored by the reference; the arithmetic below restates OpenCV 4.x's 8-bit fixed-point paths and is
// pinned bit-exactly to cv2 4.13 through oracle/cvprims.cpp (tests/test_oracle_cvprims.py, tests/test_gpu_intake.py):
//   gray  = (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   remap : map coordinates rounded to 1/32 px (cvRound(x*32), round-half-even), bilinear weights (32-fx)(32-fy)*32 ...
//           (sum 2^15), result (sum + 2^14) >> 15, taps outside the source = 0 (BORDER_CONSTANT)
//   resize: INTER_LINEAR with 11-bit coefficients (2x2 decimation is OpenCV's area path), the same arithmetic as the
//           pyramid levels (pyramid.cu)
// All three are HBM-streaming kernels (1 output byte per 1..4 input bytes, one output pixel per thread).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "orbfe_internal.h"
#include "remap_core.h"
#include "scratch.h"

namespace {

int ifail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define ICK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return ifail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

// channels = 3 or 4; rgb != 0: the first channel is R (COLOR_RGB2GRAY / RGBA2GRAY), else B.
__global__ void __launch_bounds__(256)
k_cvt_gray(const uint8_t* __restrict__ src, int rows, int cols, size_t sstep, int channels, int rgb,
           uint8_t* __restrict__ dst, size_t dstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols || y >= rows) return;
    const uint8_t* p = src + (size_t)y * sstep + (size_t)x * channels;
    const int c0 = p[0], c1 = p[1], c2 = p[2];
    const int b = rgb ? c2 : c0, r = rgb ? c0 : c2;
    dst[(size_t)y * dstep + x] = (uint8_t)((b * 3735 + c1 * 19235 + r * 9798 + (1 << 14)) >> 15);
}

__global__ void __launch_bounds__(256)
k_remap_linear(const uint8_t* __restrict__ src, int srows, int scols, size_t sstep, const float* __restrict__ mapx,
               const float* __restrict__ mapy, size_t mstep, int drows, int dcols, uint8_t* __restrict__ dst, size_t dstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dcols || y >= drows) return;
    dst[(size_t)y * dstep + x] =
        (uint8_t)orbfe_remap_sample(src, sstep, srows, scols, mapx[(size_t)y * mstep + x], mapy[(size_t)y * mstep + x]);
}

struct ITap { int s0, s1; short a0, a1; };

__global__ void __launch_bounds__(256)
k_resize_linear(const uint8_t* __restrict__ src, size_t sstep, const ITap* __restrict__ xt, const ITap* __restrict__ yt,
                int drows, int dcols, uint8_t* __restrict__ dst, size_t dstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dcols || y >= drows) return;
    const ITap tx = xt[x], ty = yt[y];
    const uint8_t* s0 = src + (size_t)ty.s0 * sstep;
    const uint8_t* s1 = src + (size_t)ty.s1 * sstep;
    const int h0 = s0[tx.s0] * tx.a0 + s0[tx.s1] * tx.a1, h1 = s1[tx.s0] * tx.a0 + s1[tx.s1] * tx.a1;
    dst[(size_t)y * dstep + x] = (uint8_t)((((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

__global__ void __launch_bounds__(256)
k_resize_area2(const uint8_t* __restrict__ src, size_t sstep, int drows, int dcols, uint8_t* __restrict__ dst, size_t dstep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dcols || y >= drows) return;
    const uint8_t* s0 = src + (size_t)(2 * y) * sstep + 2 * x;
    const uint8_t* s1 = s0 + sstep;
    dst[(size_t)y * dstep + x] = (uint8_t)((s0[0] + s0[1] + s1[0] + s1[1] + 2) >> 2);
}

// Frame::UndistortKeyPoints (src/Frame.cc:1003-1051) = cv::undistortPoints(pts, K, mDistCoef, noArray(), mK): five
// fixed-point iterations of the distortion model in double, re-projection with K, result stored as float.  Compiled
// with --fmad=false: every product and sum rounds like OpenCV's scalar code.
struct UndistPrm { double fx, fy, cx, cy, k[14]; };

__global__ void k_undistort(const OrbfeKeyPoint* __restrict__ in, int n, UndistPrm P, OrbfeKeyPoint* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OrbfeKeyPoint kp = in[i];
    const double ifx = 1. / P.fx, ify = 1. / P.fy;
    const double u = kp.x, v = kp.y;
    double x = (u - P.cx) * ifx, y = (v - P.cy) * ify;
    const double x0 = x, y0 = y;
    const double* k = P.k;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = (u - P.cx) * ifx; y = (v - P.cy) * ify; break; }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = P.fx * x + 0 * y + P.cx, yy = 0 * x + P.fy * y + P.cy, ww = 1. / (0 * x + 0 * y + 1);
    kp.x = (float)(xx * ww);
    kp.y = (float)(yy * ww);
    out[i] = kp;
}

inline int cv_floor_f(float v) { return (int)floorf(v); }
inline int cv_round_f(float v) { return (int)lrintf(v); }

// OpenCV resize INTER_LINEAR coefficient tables (same arithmetic as build_taps in orbfe_api.cu)
void build_itaps(int ssize, int dsize, bool isX, std::vector<ITap>& out) {
    out.resize(dsize);
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor_f(f);
        f -= s;
        ITap t;
        if (isX) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
            t.s0 = s; t.s1 = std::min(s + 1, ssize - 1);
        } else {
            t.s0 = std::min(std::max(s, 0), ssize - 1);
            t.s1 = std::min(std::max(s + 1, 0), ssize - 1);
        }
        t.a0 = (short)cv_round_f((1.f - f) * 2048.f);
        t.a1 = (short)cv_round_f(f * 2048.f);
        out[d] = t;
    }
}

int check_dev(int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return ifail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return ifail(ORBFE_ERR_INVALID, "bad device ordinal");
    return ORBFE_OK;
}

}  // namespace

// ---- device-pointer entry points (no synchronisation) -------------------------------------------------------
extern "C" int orbfe_cvt_gray_device(const uint8_t* d_src, int rows, int cols, size_t src_step, int channels, int rgb_order,
                                     uint8_t* d_dst, size_t dst_step, void* stream) {
    if (!d_src || !d_dst || rows <= 0 || cols <= 0 || (channels != 3 && channels != 4) || src_step < (size_t)cols * channels ||
        dst_step < (size_t)cols)
        return ifail(ORBFE_ERR_INVALID, "cvtColor: bad arguments");
    k_cvt_gray<<<dim3((cols + 255) / 256, rows), 256, 0, (cudaStream_t)stream>>>(d_src, rows, cols, src_step, channels,
                                                                                   rgb_order ? 1 : 0, d_dst, dst_step);
    ICK(cudaGetLastError());
    return ORBFE_OK;
}

extern "C" int orbfe_remap_linear_device(const uint8_t* d_src, int src_rows, int src_cols, size_t src_step,
                                         const float* d_map_x, const float* d_map_y, size_t map_step_floats, int dst_rows,
                                         int dst_cols, uint8_t* d_dst, size_t dst_step, void* stream) {
    if (!d_src || !d_dst || !d_map_x || !d_map_y || src_rows <= 0 || src_cols <= 0 || dst_rows <= 0 || dst_cols <= 0 ||
        src_step < (size_t)src_cols || dst_step < (size_t)dst_cols || map_step_floats < (size_t)dst_cols)
        return ifail(ORBFE_ERR_INVALID, "remap: bad arguments");
    k_remap_linear<<<dim3((dst_cols + 255) / 256, dst_rows), 256, 0, (cudaStream_t)stream>>>(
        d_src, src_rows, src_cols, src_step, d_map_x, d_map_y, map_step_floats, dst_rows, dst_cols, d_dst, dst_step);
    ICK(cudaGetLastError());
    return ORBFE_OK;
}

// ---- host-pointer entry points ----------------------------------------------------------------------------------
extern "C" int orbfe_cvt_gray(const uint8_t* src, int rows, int cols, size_t src_step, int channels, int rgb_order,
                              uint8_t* dst, size_t dst_step, int device) {
    int rc = check_dev(device);
    if (rc != ORBFE_OK) return rc;
    if (!src || !dst || rows <= 0 || cols <= 0 || (channels != 3 && channels != 4) || src_step < (size_t)cols * channels ||
        dst_step < (size_t)cols)
        return ifail(ORBFE_ERR_INVALID, "cvtColor: bad arguments");
    OrbfeStage S;
    const size_t iS = S.in(src, src_step * (size_t)(rows - 1) + (size_t)cols * channels);
    const size_t oD = S.out(nullptr, (size_t)rows * cols);
    ICK(S.commit(device));
    ICK(S.upload());
    rc = orbfe_cvt_gray_device(S.ptr<uint8_t>(iS), rows, cols, src_step, channels, rgb_order, S.ptr<uint8_t>(oD), (size_t)cols,
                               S.stream());
    if (rc != ORBFE_OK) return rc;
    ICK(cudaMemcpy2DAsync(dst, dst_step, S.ptr<uint8_t>(oD), (size_t)cols, (size_t)cols, rows, cudaMemcpyDeviceToHost, S.stream()));
    ICK(cudaStreamSynchronize(S.stream()));
    return ORBFE_OK;
}

extern "C" int orbfe_remap_linear(const uint8_t* src, int src_rows, int src_cols, size_t src_step, const float* map_x,
                                  const float* map_y, int dst_rows, int dst_cols, uint8_t* dst, size_t dst_step, int device) {
    int rc = check_dev(device);
    if (rc != ORBFE_OK) return rc;
    if (!src || !dst || !map_x || !map_y || src_rows <= 0 || src_cols <= 0 || dst_rows <= 0 || dst_cols <= 0 ||
        src_step < (size_t)src_cols || dst_step < (size_t)dst_cols)
        return ifail(ORBFE_ERR_INVALID, "remap: bad arguments");
    OrbfeStage S;
    const size_t nmap = 4 * (size_t)dst_rows * dst_cols;
    const size_t iS = S.in(src, src_step * (size_t)(src_rows - 1) + (size_t)src_cols);
    const size_t iX = S.in(map_x, nmap), iY = S.in(map_y, nmap);
    const size_t oD = S.out(nullptr, (size_t)dst_rows * dst_cols);
    ICK(S.commit(device));
    ICK(S.upload());
    rc = orbfe_remap_linear_device(S.ptr<uint8_t>(iS), src_rows, src_cols, src_step, S.ptr<float>(iX), S.ptr<float>(iY),
                                   (size_t)dst_cols, dst_rows, dst_cols, S.ptr<uint8_t>(oD), (size_t)dst_cols, S.stream());
    if (rc != ORBFE_OK) return rc;
    ICK(cudaMemcpy2DAsync(dst, dst_step, S.ptr<uint8_t>(oD), (size_t)dst_cols, (size_t)dst_cols, dst_rows,
                          cudaMemcpyDeviceToHost, S.stream()));
    ICK(cudaStreamSynchronize(S.stream()));
    return ORBFE_OK;
}

extern "C" int orbfe_resize_linear(const uint8_t* src, int src_rows, int src_cols, size_t src_step, int dst_rows,
                                   int dst_cols, uint8_t* dst, size_t dst_step, int device) {
    int rc = check_dev(device);
    if (rc != ORBFE_OK) return rc;
    if (!src || !dst || src_rows <= 0 || src_cols <= 0 || dst_rows <= 0 || dst_cols <= 0 || src_step < (size_t)src_cols ||
        dst_step < (size_t)dst_cols)
        return ifail(ORBFE_ERR_INVALID, "resize: bad arguments");
    std::vector<ITap> xt, yt;
    const double sx = 1.0 / ((double)dst_cols / src_cols), sy = 1.0 / ((double)dst_rows / src_rows);
    const bool same = dst_cols == src_cols && dst_rows == src_rows;
    const bool area2 = !same && std::abs(sx - 2.0) < DBL_EPSILON && std::abs(sy - 2.0) < DBL_EPSILON;
    if (!area2) { build_itaps(src_cols, dst_cols, true, xt); build_itaps(src_rows, dst_rows, false, yt); }
    OrbfeStage S;
    const size_t iS = S.in(src, src_step * (size_t)(src_rows - 1) + (size_t)src_cols);
    const size_t iX = S.in(xt.data(), sizeof(ITap) * xt.size()), iY = S.in(yt.data(), sizeof(ITap) * yt.size());
    const size_t oD = S.out(nullptr, (size_t)dst_rows * dst_cols);
    ICK(S.commit(device));
    ICK(S.upload());
    cudaStream_t st = S.stream();
    const dim3 grid((dst_cols + 255) / 256, dst_rows);
    if (same)
        ICK(cudaMemcpy2DAsync(S.ptr<uint8_t>(oD), (size_t)dst_cols, S.ptr<uint8_t>(iS), src_step, (size_t)dst_cols, dst_rows,
                              cudaMemcpyDeviceToDevice, st));
    else if (area2)
        k_resize_area2<<<grid, 256, 0, st>>>(S.ptr<uint8_t>(iS), src_step, dst_rows, dst_cols, S.ptr<uint8_t>(oD), (size_t)dst_cols);
    else
        k_resize_linear<<<grid, 256, 0, st>>>(S.ptr<uint8_t>(iS), src_step, S.ptr<ITap>(iX), S.ptr<ITap>(iY), dst_rows, dst_cols,
                                              S.ptr<uint8_t>(oD), (size_t)dst_cols);
    ICK(cudaGetLastError());
    ICK(cudaMemcpy2DAsync(dst, dst_step, S.ptr<uint8_t>(oD), (size_t)dst_cols, (size_t)dst_cols, dst_rows,
                          cudaMemcpyDeviceToHost, st));
    ICK(cudaStreamSynchronize(st));
    return ORBFE_OK;
}

extern "C" int orbfe_undistort_keypoints(const OrbfeKeyPoint* keys, int n, float fx, float fy, float cx, float cy,
                                         const float* dist_coef, int n_dist, OrbfeKeyPoint* keys_un, int device) {
    if (n < 0 || (n > 0 && (!keys || !keys_un)) || n_dist < 0 || n_dist > 14 || (n_dist > 0 && !dist_coef))
        return ifail(ORBFE_ERR_INVALID, "undistort: bad arguments");
    if (n == 0) return ORBFE_OK;
    if (n_dist == 0 || dist_coef[0] == 0.0f) {   // Frame.cc:1005-1009: mvKeysUn = mvKeys
        if (keys_un != keys) memcpy(keys_un, keys, sizeof(OrbfeKeyPoint) * (size_t)n);
        return ORBFE_OK;
    }
    int rc = check_dev(device);
    if (rc != ORBFE_OK) return rc;
    UndistPrm P;
    P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
    for (int i = 0; i < 14; i++) P.k[i] = i < n_dist ? (double)dist_coef[i] : 0.0;
    OrbfeStage S;
    const size_t iK = S.in(keys, sizeof(OrbfeKeyPoint) * (size_t)n), oK = S.out(keys_un, sizeof(OrbfeKeyPoint) * (size_t)n);
    ICK(S.commit(device));
    ICK(S.upload());
    k_undistort<<<(n + 127) / 128, 128, 0, S.stream()>>>(S.ptr<OrbfeKeyPoint>(iK), n, P, S.ptr<OrbfeKeyPoint>(oK));
    ICK(cudaGetLastError());
    ICK(S.download());
    return ORBFE_OK;
}
