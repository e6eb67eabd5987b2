// kb8_core.h -- the per-point / per-match arithmetic of csrc/kb8.cu (KannalaBrandt8::project, ::unproject,
// ::TriangulateMatches + ::Triangulate; reference lines cited in kb8.cu).  Like octree_core.h and fast_core.h the same
// source compiles for the host, so the product logic is unit-tested on the CPU against the oracle
// (tests/native/kb8_core_host.cpp, tests/test_kb8_core.py); the kernels in kb8.cu are thin loops around it.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#else
#define KB_HD inline
#endif

struct Kb8Cam { float p[8]; float precision; };
struct Kb8Rig { Kb8Cam c1, c2; float R12[9]; float t12[3]; };

static KB_HD float f_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }

static KB_HD void kb8_project(const Kb8Cam& c, const float* p3, float* uv) {
    const float x2_plus_y2 = p3[0] * p3[0] + p3[1] * p3[1];
    const float theta = f_atan2(sqrtf(x2_plus_y2), p3[2]);
    const float psi = f_atan2(p3[1], p3[0]);
    const float theta2 = theta * theta;
    const float theta3 = theta * theta2;
    const float theta5 = theta3 * theta2;
    const float theta7 = theta5 * theta2;
    const float theta9 = theta7 * theta2;
    const float r = theta + c.p[4] * theta3 + c.p[5] * theta5 + c.p[6] * theta7 + c.p[7] * theta9;
    // the reference's unqualified cos(psi) / sin(psi) are the double functions: product and sum in double, one rounding
    uv[0] = (float)((double)(c.p[0] * r) * cos((double)psi) + (double)c.p[2]);
    uv[1] = (float)((double)(c.p[1] * r) * sin((double)psi) + (double)c.p[3]);
}

static KB_HD void kb8_unproject(const Kb8Cam& c, const float* uv, float* ray) {
    const float pwx = (uv[0] - c.p[2]) / c.p[0], pwy = (uv[1] - c.p[3]) / c.p[1];
    float scale = 1.f;
    float theta_d = sqrtf(pwx * pwx + pwy * pwy);
    const float halfPi = (float)(3.1415926535897932384626433832795 / 2.0);
    theta_d = fminf(fmaxf(-halfPi, theta_d), halfPi);
    if ((double)theta_d > 1e-8) {
        float theta = theta_d;
        for (int j = 0; j < 10; j++) {
            const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
            const float k0 = c.p[4] * theta2, k1 = c.p[5] * theta4, k2 = c.p[6] * theta6, k3 = c.p[7] * theta8;
            const float fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
            theta = theta - fix;
            if (fabsf(fix) < c.precision) break;
        }
        scale = (float)tan((double)theta) / theta_d;
    }
    ray[0] = pwx * scale;
    ray[1] = pwy * scale;
    ray[2] = 1.f;
}

// Eigenvector of A^T A for the smallest eigenvalue (= right singular vector of the smallest singular value of A).
static KB_HD void null_vector4(const float A[4][4], double x[4]) {
    double S[4][4], V[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) s += (double)A[k][i] * (double)A[k][j];
            S[i][j] = s;
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0;
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) off += S[p][q] * S[p][q];
        if (off < 1e-300) break;
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                if (S[p][q] == 0.0) continue;
                const double tau = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double a = S[k][p], b = S[k][q];
                    S[k][p] = c * a - s * b;
                    S[k][q] = s * a + c * b;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double a = S[p][k], b = S[q][k];
                    S[p][k] = c * a - s * b;
                    S[q][k] = s * a + c * b;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double a = V[k][p], b = V[k][q];
                    V[k][p] = c * a - s * b;
                    V[k][q] = s * a + c * b;
                }
            }
    }
    double best = S[0][0];
    double v0 = V[0][0], v1 = V[1][0], v2 = V[2][0], v3 = V[3][0];
#pragma unroll
    for (int i = 1; i < 4; i++)
        if (S[i][i] < best) { best = S[i][i]; v0 = V[0][i]; v1 = V[1][i]; v2 = V[2][i]; v3 = V[3][i]; }
    x[0] = v0; x[1] = v1; x[2] = v2; x[3] = v3;
}

// One TriangulateMatches call: returns the reference's return value; x3D is the triangulated point (valid, i.e. the
// reference would have written p3D, iff the return value is z1 > 0 -- `ok`).
static KB_HD float kb8_triangulate_one(const Kb8Rig& rig, const float* a1, const float* a2, float sigma1, float unc2,
                                       float* x3D, bool& ok) {
    float r1[3], r2[3];
    kb8_unproject(rig.c1, a1, r1);
    kb8_unproject(rig.c2, a2, r2);
    const float* R12 = rig.R12;
    float r21[3];
#pragma unroll
    for (int k = 0; k < 3; k++) r21[k] = R12[3 * k] * r2[0] + R12[3 * k + 1] * r2[1] + R12[3 * k + 2] * r2[2];
    const float dot = r1[0] * r21[0] + r1[1] * r21[1] + r1[2] * r21[2];
    const float n1 = sqrtf(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
    const float n21 = sqrtf(r21[0] * r21[0] + r21[1] * r21[1] + r21[2] * r21[2]);
    const float cosParallaxRays = dot / (n1 * n21);
    float result;
    x3D[0] = x3D[1] = x3D[2] = 0.f;
    ok = false;
    if ((double)cosParallaxRays > 0.9998) {
        result = -1.f;
    } else {
        float R21[3][3], T2[3][4];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) R21[r][c] = R12[3 * c + r];
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int c = 0; c < 3; c++) T2[r][c] = R21[r][c];
            T2[r][3] = (-R21[r][0]) * rig.t12[0] + (-R21[r][1]) * rig.t12[1] + (-R21[r][2]) * rig.t12[2];
        }
        const float T1[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
        float A[4][4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            A[0][c] = r1[0] * T1[2][c] - T1[0][c];
            A[1][c] = r1[1] * T1[2][c] - T1[1][c];
            A[2][c] = r2[0] * T2[2][c] - T2[0][c];
            A[3][c] = r2[1] * T2[2][c] - T2[1][c];
        }
        double xh[4];
        null_vector4(A, xh);
        x3D[0] = (float)xh[0] / (float)xh[3];
        x3D[1] = (float)xh[1] / (float)xh[3];
        x3D[2] = (float)xh[2] / (float)xh[3];
        const float z1 = x3D[2];
        const float z2 = R21[2][0] * x3D[0] + R21[2][1] * x3D[1] + R21[2][2] * x3D[2] + T2[2][3];
        if (z1 <= 0) {
            result = -2.f;
        } else if (z2 <= 0) {
            result = -3.f;
        } else {
            float uv1[2];
            kb8_project(rig.c1, x3D, uv1);
            const float ex1 = uv1[0] - a1[0], ey1 = uv1[1] - a1[1];
            if ((double)(ex1 * ex1 + ey1 * ey1) > 5.991 * (double)sigma1) {
                result = -4.f;
            } else {
                float x3D2[3];
#pragma unroll
                for (int r = 0; r < 3; r++) x3D2[r] = R21[r][0] * x3D[0] + R21[r][1] * x3D[1] + R21[r][2] * x3D[2] + T2[r][3];
                float uv2[2];
                kb8_project(rig.c2, x3D2, uv2);
                const float ex2 = uv2[0] - a2[0], ey2 = uv2[1] - a2[1];
                if ((double)(ex2 * ex2 + ey2 * ey2) > 5.991 * (double)unc2) {
                    result = -5.f;
                } else {
                    result = z1;
                    ok = true;
                }
            }
        }
    }
    return result;
}
