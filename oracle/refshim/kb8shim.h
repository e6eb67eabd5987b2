// oracle/refshim/kb8shim.h -- TEST INFRASTRUCTURE ONLY.
// Just enough of include/CameraModels/KannalaBrandt8.h (members mvParameters, precision) and of the cv / Eigen value
// types for the reference's OWN bodies of KannalaBrandt8::project (cv::Point3f and Eigen::Vector3f overloads) and
// KannalaBrandt8::unproject -- sliced at build time out of /root/reference/src/CameraModels/KannalaBrandt8.cpp by
// oracle/ref_slices.py (mode kb8), never copied into this repo -- to compile without Eigen / Boost.
#pragma once
#include <cmath>
#include <vector>

#include <opencv2/core/core.hpp>

namespace cv {
struct Point3f {
    float x, y, z;
    Point3f() : x(0), y(0), z(0) {}
    Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
}  // namespace cv

namespace Eigen {
struct Vector3f {
    float v[3];
    float operator[](int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
};
struct Vector2f {
    float v[2];
    float operator[](int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
};
}  // namespace Eigen

namespace ORB_SLAM3 {
class KannalaBrandt8 {
   public:
    KannalaBrandt8(const float* p, float prec) : mvParameters(p, p + 8), precision(prec) {}
    cv::Point2f project(const cv::Point3f& p3D);
    Eigen::Vector2f project(const Eigen::Vector3f& v3D);
    cv::Point3f unproject(const cv::Point2f& p2D);
    std::vector<float> mvParameters;   // include/CameraModels/GeometricCamera.h
    const float precision;             // include/CameraModels/KannalaBrandt8.h:102
};
}  // namespace ORB_SLAM3
