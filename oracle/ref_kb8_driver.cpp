// oracle/ref_kb8_driver.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points onto the reference's own
// KannalaBrandt8::project / unproject bodies (sliced by ref_slices.py kb8, compiled against refshim/kb8shim.h).
#include "kb8shim.h"

extern "C" {
void ref_kb8_project(const float* P, const float* p3d, int n, float* uv) {
    ORB_SLAM3::KannalaBrandt8 cam(P, 1e-6f);
    for (int i = 0; i < n; i++) {
        const cv::Point2f r = cam.project(cv::Point3f(p3d[3 * i], p3d[3 * i + 1], p3d[3 * i + 2]));
        uv[2 * i] = r.x; uv[2 * i + 1] = r.y;
    }
}
void ref_kb8_project_eig(const float* P, const float* p3d, int n, float* uv) {
    ORB_SLAM3::KannalaBrandt8 cam(P, 1e-6f);
    for (int i = 0; i < n; i++) {
        Eigen::Vector3f v;
        v[0] = p3d[3 * i]; v[1] = p3d[3 * i + 1]; v[2] = p3d[3 * i + 2];
        const Eigen::Vector2f r = cam.project(v);
        uv[2 * i] = r[0]; uv[2 * i + 1] = r[1];
    }
}
void ref_kb8_unproject(const float* P, float precision, const float* uv, int n, float* rays) {
    ORB_SLAM3::KannalaBrandt8 cam(P, precision);
    for (int i = 0; i < n; i++) {
        const cv::Point3f r = cam.unproject(cv::Point2f(uv[2 * i], uv[2 * i + 1]));
        rays[3 * i] = r.x; rays[3 * i + 1] = r.y; rays[3 * i + 2] = r.z;
    }
}
}
