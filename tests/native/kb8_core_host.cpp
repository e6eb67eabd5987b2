// tests/native/kb8_core_host.cpp -- compiles the product's per-point / per-match KB8 arithmetic
// (orb-slam3_byzyh_b200/csrc/kb8_core.h, the body of the kernels in kb8.cu) for the HOST so that its logic can be
// unit-tested on the CPU against the oracle.  A test of product logic, not a CPU path of the product.
#include <cmath>
#include <limits>

#include "../../orb-slam3_byzyh_b200/csrc/kb8_core.h"

static Kb8Cam cam(const float* p, float prec) {
    Kb8Cam c;
    for (int i = 0; i < 8; i++) c.p[i] = p[i];
    c.precision = prec;
    return c;
}

extern "C" {
void kb8_core_project(const float* P, const float* p3d, int n, float* uv) {
    const Kb8Cam c = cam(P, 1e-6f);
    for (int i = 0; i < n; i++) kb8_project(c, p3d + 3 * i, uv + 2 * i);
}
void kb8_core_unproject(const float* P, float prec, const float* uv, int n, float* rays) {
    const Kb8Cam c = cam(P, prec);
    for (int i = 0; i < n; i++) kb8_unproject(c, uv + 2 * i, rays + 3 * i);
}
void kb8_core_triangulate(const float* P1, float prec1, const float* P2, float prec2, const float* R12, const float* t12,
                          const float* pt1, const float* pt2, const float* sigma1, const float* unc2, int n, float* depth,
                          float* p3d) {
    Kb8Rig rig;
    rig.c1 = cam(P1, prec1);
    rig.c2 = cam(P2, prec2);
    for (int i = 0; i < 9; i++) rig.R12[i] = R12[i];
    for (int i = 0; i < 3; i++) rig.t12[i] = t12[i];
    for (int i = 0; i < n; i++) {
        float x[3];
        bool ok;
        depth[i] = kb8_triangulate_one(rig, pt1 + 2 * i, pt2 + 2 * i, sigma1[i], unc2[i], x, ok);
        for (int k = 0; k < 3; k++) p3d[3 * i + k] = ok ? x[k] : std::numeric_limits<float>::quiet_NaN();
    }
}
}
