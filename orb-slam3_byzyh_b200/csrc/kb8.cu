// kb8.cu -- KannalaBrandt8 geometry behind the fisheye stereo matcher, batched.
//   k_kb8_project      KannalaBrandt8::project      /root/reference/src/CameraModels/KannalaBrandt8.cpp:40-55, 84-101
//   k_kb8_unproject    KannalaBrandt8::unproject    :180-217 (Newton on theta, at most 10 steps, `precision` stop)
//   k_kb8_triangulate  KannalaBrandt8::TriangulateMatches :439-515 + ::Triangulate :553-565, i.e. the per-match body of
//                      Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1560-1587) and ::epipolarConstrain (:322-328)
// One thread per point / match.  fp32 in the reference's operation order (the library is built with --fmad=false);
// the transcendental calls (atan2f, tan, and the double cos / sin the reference's unqualified calls resolve to) are
// evaluated in double and rounded once, which reproduces glibc's results except where those are not correctly rounded.
// The 4x4 DLT null vector (Eigen::JacobiSVD in the reference) is the eigenvector of A^T A for its smallest eigenvalue,
// cyclic Jacobi in fp64: 4x4 per thread, no library call.
#include <math.h>
#include <stdint.h>

#include "kb8_core.h"
#include "orbfe_internal.h"
#include "scratch.h"

namespace {

int kfail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define KCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return kfail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

__global__ void k_kb8_project(Kb8Cam c, const float* __restrict__ p3d, int n, float* __restrict__ uv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p[3] = {p3d[3 * i], p3d[3 * i + 1], p3d[3 * i + 2]};
    float o[2];
    kb8_project(c, p, o);
    uv[2 * i] = o[0];
    uv[2 * i + 1] = o[1];
}

__global__ void k_kb8_unproject(Kb8Cam c, const float* __restrict__ uv, int n, float* __restrict__ rays) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p[2] = {uv[2 * i], uv[2 * i + 1]};
    float r[3];
    kb8_unproject(c, p, r);
    rays[3 * i] = r[0];
    rays[3 * i + 1] = r[1];
    rays[3 * i + 2] = r[2];
}

__global__ void k_kb8_triangulate(const __grid_constant__ Kb8Rig rig, const float* __restrict__ pt1,
                                  const float* __restrict__ pt2, const float* __restrict__ sigma1,
                                  const float* __restrict__ unc2, int n, float* __restrict__ depth,
                                  float* __restrict__ p3d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a1[2] = {pt1[2 * i], pt1[2 * i + 1]}, a2[2] = {pt2[2 * i], pt2[2 * i + 1]};
    float x3D[3];
    bool ok;
    const float result = kb8_triangulate_one(rig, a1, a2, sigma1[i], unc2[i], x3D, ok);
    depth[i] = result;
    if (ok) {   // the reference leaves p3D untouched on rejection; here rejected rows are NaN
        p3d[3 * i] = x3D[0]; p3d[3 * i + 1] = x3D[1]; p3d[3 * i + 2] = x3D[2];
    } else {
        const float nanv = __int_as_float(0x7fc00000);
        p3d[3 * i] = nanv; p3d[3 * i + 1] = nanv; p3d[3 * i + 2] = nanv;
    }
}

int kb8_check_dev(int device) {
    int ndev = 0;
    const cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return kfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return kfail(ORBFE_ERR_INVALID, "bad device ordinal");
    return ORBFE_OK;
}

Kb8Cam make_cam(const float* params, float precision) {
    Kb8Cam c;
    for (int i = 0; i < 8; i++) c.p[i] = params[i];
    c.precision = precision;
    return c;
}

}  // namespace

extern "C" int orbfe_kb8_project(const float* params, const float* p3d, int n, float* uv, int device) {
    if (!params || n < 0 || (n > 0 && (!p3d || !uv))) return kfail(ORBFE_ERR_INVALID, "kb8_project: bad arguments");
    int rc = kb8_check_dev(device);
    if (rc != ORBFE_OK) return rc;
    if (n == 0) return ORBFE_OK;
    OrbfeStage S;
    const size_t iP = S.in(p3d, sizeof(float) * 3 * (size_t)n), oU = S.out(uv, sizeof(float) * 2 * (size_t)n);
    KCK(S.commit(device));
    KCK(S.upload());
    k_kb8_project<<<(n + 127) / 128, 128, 0, S.stream()>>>(make_cam(params, 1e-6f), S.ptr<float>(iP), n, S.ptr<float>(oU));
    KCK(cudaGetLastError());
    KCK(S.download());
    return ORBFE_OK;
}

extern "C" int orbfe_kb8_unproject(const float* params, float precision, const float* uv, int n, float* rays, int device) {
    if (!params || n < 0 || (n > 0 && (!uv || !rays))) return kfail(ORBFE_ERR_INVALID, "kb8_unproject: bad arguments");
    int rc = kb8_check_dev(device);
    if (rc != ORBFE_OK) return rc;
    if (n == 0) return ORBFE_OK;
    OrbfeStage S;
    const size_t iU = S.in(uv, sizeof(float) * 2 * (size_t)n), oR = S.out(rays, sizeof(float) * 3 * (size_t)n);
    KCK(S.commit(device));
    KCK(S.upload());
    k_kb8_unproject<<<(n + 127) / 128, 128, 0, S.stream()>>>(make_cam(params, precision), S.ptr<float>(iU), n, S.ptr<float>(oR));
    KCK(cudaGetLastError());
    KCK(S.download());
    return ORBFE_OK;
}

// The step that stands where the reference calls Eigen::JacobiSVD<Matrix4f> (KannalaBrandt8.cpp:566-568), on its own: the
// homogeneous solution of n 4 x 4 systems.  A stage tap: the oracle restates the same fp64 cyclic Jacobi, and with
// IEEE add / mul / div / sqrt and no contraction on either side the two must agree bit for bit (tests/test_gpu_kb8.py).
__global__ void k_kb8_null_vectors(const float* __restrict__ A, int n, double* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float M[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) M[r][c] = A[16 * (size_t)i + 4 * r + c];
    double v[4];
    null_vector4(M, v);
#pragma unroll
    for (int k = 0; k < 4; k++) x[4 * (size_t)i + k] = v[k];
}

extern "C" int orbfe_debug_kb8_null_vectors(const float* A, int n, double* x, int device) {
    if (n < 0 || (n > 0 && (!A || !x))) return kfail(ORBFE_ERR_INVALID, "kb8_null_vectors: bad arguments");
    int rc = kb8_check_dev(device);
    if (rc != ORBFE_OK) return rc;
    if (n == 0) return ORBFE_OK;
    OrbfeStage S;
    const size_t iA = S.in(A, sizeof(float) * 16 * (size_t)n), oX = S.out(x, sizeof(double) * 4 * (size_t)n);
    KCK(S.commit(device));
    KCK(S.upload());
    k_kb8_null_vectors<<<(n + 63) / 64, 64, 0, S.stream()>>>(S.ptr<float>(iA), n, S.ptr<double>(oX));
    KCK(cudaGetLastError());
    KCK(S.download());
    return ORBFE_OK;
}

extern "C" int orbfe_kb8_triangulate_matches(const float* params1, float precision1, const float* params2,
                                             float precision2, const float* R12, const float* t12, const float* pt1,
                                             const float* pt2, const float* sigma1, const float* unc2, int n,
                                             float* depth, float* p3d, int device) {
    if (!params1 || !params2 || !R12 || !t12 || n < 0 || (n > 0 && (!pt1 || !pt2 || !sigma1 || !unc2 || !depth || !p3d)))
        return kfail(ORBFE_ERR_INVALID, "kb8_triangulate_matches: bad arguments");
    int rc = kb8_check_dev(device);
    if (rc != ORBFE_OK) return rc;
    if (n == 0) return ORBFE_OK;
    Kb8Rig rig;
    rig.c1 = make_cam(params1, precision1);
    rig.c2 = make_cam(params2, precision2);
    for (int i = 0; i < 9; i++) rig.R12[i] = R12[i];
    for (int i = 0; i < 3; i++) rig.t12[i] = t12[i];
    OrbfeStage S;
    const size_t i1 = S.in(pt1, sizeof(float) * 2 * (size_t)n), i2 = S.in(pt2, sizeof(float) * 2 * (size_t)n);
    const size_t iS = S.in(sigma1, sizeof(float) * (size_t)n), iU = S.in(unc2, sizeof(float) * (size_t)n);
    const size_t oD = S.out(depth, sizeof(float) * (size_t)n), oP = S.out(p3d, sizeof(float) * 3 * (size_t)n);
    KCK(S.commit(device));
    KCK(S.upload());
    k_kb8_triangulate<<<(n + 63) / 64, 64, 0, S.stream()>>>(rig, S.ptr<float>(i1), S.ptr<float>(i2), S.ptr<float>(iS),
                                                           S.ptr<float>(iU), n, S.ptr<float>(oD), S.ptr<float>(oP));
    KCK(cudaGetLastError());
    KCK(S.download());
    return ORBFE_OK;
}
