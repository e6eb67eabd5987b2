// oracle/kb8_oracle.cpp -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
//
// CPU restatement of the fisheye-stereo geometry that follows the kNN matcher in
// Frame::ComputeStereoFishEyeMatches (/root/reference/src/Frame.cc:1560-1587):
//   KannalaBrandt8::project(cv::Point3f / Eigen::Vector3f)  src/CameraModels/KannalaBrandt8.cpp:40-55, 84-101
//   KannalaBrandt8::unproject(cv::Point2f)                   :180-217  (Newton on theta, at most 10 steps)
//   KannalaBrandt8::TriangulateMatches                       :439-515  (return codes -1 .. -5, else z1)
//   KannalaBrandt8::Triangulate                              :553-565  (null vector of the 4x4 DLT system)
//   KannalaBrandt8::epipolarConstrain                        :322-328  (TriangulateMatches(...) > 0.0001f)
// fp32 throughout, operations in the reference's order (compiled with -ffp-contract=off).
// Parity status: project / unproject are PINNED against the reference's own function bodies compiled verbatim
// (oracle/_ref/libref_kb8.so: ref_build.sh + ref_slices.py kb8, tests/test_oracle_kb8.py).  Triangulate is
// "parity unpinned": the reference calls Eigen::JacobiSVD<Matrix4f> and Eigen is not in this image; the null vector is
// computed here in fp64 (cyclic Jacobi on A^T A), which agrees with any backward-stable fp32 SVD to the conditioning
// of the system, so comparisons through it are tolerance-based and the return code is only compared away from its
// thresholds.
#include <cmath>
#include <cstdint>

namespace kb8_oracle {

static const float kPi = 3.1415926535897932384626433832795f;   // CV_PI as the float expressions see it

void project(const float* P, const float* p3, float* uv) {   // :84-101 (== :40-55 on the same floats)
    const float x2_plus_y2 = p3[0] * p3[0] + p3[1] * p3[1];
    const float theta = atan2f(sqrtf(x2_plus_y2), p3[2]);
    const float psi = atan2f(p3[1], p3[0]);
    const float theta2 = theta * theta;
    const float theta3 = theta * theta2;
    const float theta5 = theta3 * theta2;
    const float theta7 = theta5 * theta2;
    const float theta9 = theta7 * theta2;
    const float r = theta + P[4] * theta3 + P[5] * theta5 + P[6] * theta7 + P[7] * theta9;
    // `cos(psi)` is an unqualified call in a file without `using namespace std`: it resolves to ::cos(double), so the
    // product and the sum are evaluated in double and rounded once (this is what the verbatim build does; a build
    // where <math.h>'s C++ overloads are visible would take cosf and differ by at most 1 ulp)
    uv[0] = (float)((double)(P[0] * r) * cos((double)psi) + (double)P[2]);
    uv[1] = (float)((double)(P[1] * r) * sin((double)psi) + (double)P[3]);
}

void unproject(const float* P, float precision, const float* uv, float* ray) {   // :180-217
    const float pwx = (uv[0] - P[2]) / P[0], pwy = (uv[1] - P[3]) / P[1];
    float scale = 1.f;
    float theta_d = sqrtf(pwx * pwx + pwy * pwy);
    // fminf(fmaxf(-CV_PI / 2.f, theta_d), CV_PI / 2.f): CV_PI is a double constant, the bounds are doubles
    // converted to float at the call
    theta_d = fminf(fmaxf((float)(-3.1415926535897932384626433832795 / 2.f), theta_d),
                    (float)(3.1415926535897932384626433832795 / 2.f));
    if (theta_d > 1e-8) {
        float theta = theta_d;
        for (int j = 0; j < 10; j++) {
            float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
            float k0_theta2 = P[4] * theta2, k1_theta4 = P[5] * theta4;
            float k2_theta6 = P[6] * theta6, k3_theta8 = P[7] * theta8;
            float theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                              (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
            theta = theta - theta_fix;
            if (fabsf(theta_fix) < precision) break;
        }
        scale = std::tan(theta) / theta_d;
    }
    ray[0] = pwx * scale;
    ray[1] = pwy * scale;
    ray[2] = 1.f;
}

// Null vector (right singular vector of the smallest singular value) of a 4x4 matrix: eigenvector of A^T A for its
// smallest eigenvalue, cyclic Jacobi in fp64.
void null_vector4(const float A[4][4], double x[4]) {
    double S[4][4], V[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += (double)A[k][i] * (double)A[k][j];
            S[i][j] = s;
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) off += S[p][q] * S[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) {
                if (S[p][q] == 0.0) continue;
                const double tau = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < 4; k++) {
                    const double a = S[k][p], b = S[k][q];
                    S[k][p] = c * a - s * b;
                    S[k][q] = s * a + c * b;
                }
                for (int k = 0; k < 4; k++) {
                    const double a = S[p][k], b = S[q][k];
                    S[p][k] = c * a - s * b;
                    S[q][k] = s * a + c * b;
                }
                for (int k = 0; k < 4; k++) {
                    const double a = V[k][p], b = V[k][q];
                    V[k][p] = c * a - s * b;
                    V[k][q] = s * a + c * b;
                }
            }
    }
    int m = 0;
    for (int i = 1; i < 4; i++)
        if (S[i][i] < S[m][m]) m = i;
    for (int k = 0; k < 4; k++) x[k] = V[k][m];
}

// TriangulateMatches (:439-515).  R12 row-major, pt = keypoint .pt.  Returns the reference's return value; p3D is
// written only on success (as the reference leaves it untouched otherwise).
float triangulate_matches(const float* P1, float prec1, const float* P2, float prec2, const float* R12, const float* t12,
                          const float* pt1, const float* pt2, float sigmaLevel, float unc, float* p3D) {
    float r1[3], r2[3];
    unproject(P1, prec1, pt1, r1);
    unproject(P2, prec2, pt2, r2);
    float r21[3];
    for (int i = 0; i < 3; i++) r21[i] = R12[3 * i] * r2[0] + R12[3 * i + 1] * r2[1] + R12[3 * i + 2] * r2[2];
    const float dot = r1[0] * r21[0] + r1[1] * r21[1] + r1[2] * r21[2];
    const float n1 = std::sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
    const float n21 = std::sqrt(r21[0] * r21[0] + r21[1] * r21[1] + r21[2] * r21[2]);
    const float cosParallaxRays = dot / (n1 * n21);
    if (cosParallaxRays > 0.9998) return -1;
    // Tcw1 = [I | 0], Tcw2 = [R21 | -R21 t12]
    float R21[3][3], T2[3][4];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R21[i][j] = R12[3 * j + i];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T2[i][j] = R21[i][j];
        T2[i][3] = (-R21[i][0]) * t12[0] + (-R21[i][1]) * t12[1] + (-R21[i][2]) * t12[2];
    }
    const float T1[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
    float A[4][4];
    for (int j = 0; j < 4; j++) {   // Triangulate (:553-565)
        A[0][j] = r1[0] * T1[2][j] - T1[0][j];
        A[1][j] = r1[1] * T1[2][j] - T1[1][j];
        A[2][j] = r2[0] * T2[2][j] - T2[0][j];
        A[3][j] = r2[1] * T2[2][j] - T2[1][j];
    }
    double xh[4];
    null_vector4(A, xh);
    const float x3D[3] = {(float)xh[0] / (float)xh[3], (float)xh[1] / (float)xh[3], (float)xh[2] / (float)xh[3]};
    const float z1 = x3D[2];
    if (z1 <= 0) return -2;
    const float z2 = R21[2][0] * x3D[0] + R21[2][1] * x3D[1] + R21[2][2] * x3D[2] + T2[2][3];
    if (z2 <= 0) return -3;
    float uv1[2];
    project(P1, x3D, uv1);
    const float errX1 = uv1[0] - pt1[0], errY1 = uv1[1] - pt1[1];
    if ((errX1 * errX1 + errY1 * errY1) > 5.991 * sigmaLevel) return -4;
    float x3D2[3];
    for (int i = 0; i < 3; i++) x3D2[i] = R21[i][0] * x3D[0] + R21[i][1] * x3D[1] + R21[i][2] * x3D[2] + T2[i][3];
    float uv2[2];
    project(P2, x3D2, uv2);
    const float errX2 = uv2[0] - pt2[0], errY2 = uv2[1] - pt2[1];
    if ((errX2 * errX2 + errY2 * errY2) > 5.991 * unc) return -5;
    p3D[0] = x3D[0]; p3D[1] = x3D[1]; p3D[2] = x3D[2];
    return z1;
}

}  // namespace kb8_oracle

extern "C" {
void oracle_kb8_project(const float* P, const float* p3d, int n, float* uv) {
    for (int i = 0; i < n; i++) kb8_oracle::project(P, p3d + 3 * i, uv + 2 * i);
}
void oracle_kb8_unproject(const float* P, float precision, const float* uv, int n, float* rays) {
    for (int i = 0; i < n; i++) kb8_oracle::unproject(P, precision, uv + 2 * i, rays + 3 * i);
}
void oracle_kb8_null_vectors(const float* A, int n, double* x) {
    for (int i = 0; i < n; i++) {
        float M[4][4];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) M[r][c] = A[16 * (size_t)i + 4 * r + c];
        kb8_oracle::null_vector4(M, x + 4 * (size_t)i);
    }
}
void oracle_kb8_triangulate(const float* P1, float prec1, const float* P2, float prec2, const float* R12,
                            const float* t12, const float* pt1, const float* pt2, const float* sigma1,
                            const float* unc2, int n, float* depth, float* p3d) {
    for (int i = 0; i < n; i++)
        depth[i] = kb8_oracle::triangulate_matches(P1, prec1, P2, prec2, R12, t12, pt1 + 2 * i, pt2 + 2 * i, sigma1[i],
                                                   unc2[i], p3d + 3 * i);
}
}
