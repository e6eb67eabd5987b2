// oracle/cvprims.h -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
//
// CPU restatement of the OpenCV primitives that the reference's ORB front-end calls but
// does not vendor (reference CMakeLists.txt:33-39 `find_package(OpenCV 4.0)`; pinned here
// to the behaviour of OpenCV 4.13.0, the cv2 wheel in this image, and checked bit-exact
// against it by tests/test_oracle_cvprims.py):
//   resize(INTER_LINEAR, u8)          reference call site src/ORBextractor.cc:1702
//   copyMakeBorder(BORDER_REFLECT_101) src/ORBextractor.cc:1712,1734
//   FAST(type 9_16, nms)              src/ORBextractor.cc:1135,1144
//   GaussianBlur(7x7, sigma 2, u8)    src/ORBextractor.cc:1632
//   fastAtan2                         src/ORBextractor.cc:137
//   cvRound / cvFloor / cvCeil        src/ORBextractor.cc:97,160,168,519,547,553,559,1692
//   BFMatcher(NORM_HAMMING).knnMatch  src/Frame.cc:47,1553
// Everything is plain scalar C++; compile with -ffp-contract=off so fp32 expressions are
// evaluated exactly as written (no FMA contraction).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cvp {

// cvRound: round-half-to-even (SSE cvtss2si / cvtsd2si under the default MXCSR mode).
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

// BORDER_REFLECT_101 index map: gfedcb|abcdefgh|gfedcba
inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

// u8 single-channel bilinear resize, OpenCV's fixed-point INTER_LINEAR path.
void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                      uint8_t* dst, int dw, int dh, size_t dstep);

// dst is (h+top+bottom) x (w+left+right); interior copied, border = reflect-101 of src.
void copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstep,
                                 uint8_t* dst, size_t dstep, int top, int bottom, int left,
                                 int right);

struct FastKP {
    int x, y;   // column, row inside the sub-image handed to FAST
    int score;  // KeyPoint::response (integer valued)
};

// FAST-9/16 corner score of the pixel at p (ring offsets precomputed for `step`):
// max over the 16 arcs of 9 contiguous ring pixels of min(|centre - ring|) for the
// all-darker / all-brighter cases.  Pixel is a corner at threshold t iff value > t.
int fast_arc_best(const uint8_t* p, const int* ring16);
void fast_ring_offsets(int step, int* ring16);

// cv::FAST(img, kps, threshold, nonmaxSuppression, TYPE_9_16) on a w x h view.
void fast9_16(const uint8_t* img, int w, int h, size_t step, int threshold, bool nms,
              std::vector<FastKP>& out);

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1.
void gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst,
                          size_t dstep);

// cv::fastAtan2 (degrees in [0,360)).
float fast_atan2(float y, float x);

// 256-bit Hamming distance between two 32-byte descriptors.
int hamming256(const uint8_t* a, const uint8_t* b);

// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2): per query the two nearest train rows,
// ascending distance, ties -> lower train index.  idx/dist are nq x 2; missing = -1.
void bf_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist);

// cv::cvtColor(.., COLOR_{BGR,RGB,BGRA,RGBA}2GRAY), 8-bit: (B*3735 + G*19235 + R*9798 + 2^14) >> 15 (OpenCV 4.x).
void cvt_gray_u8(const uint8_t* src, int w, int h, size_t sstep, int channels, bool rgb, uint8_t* dst, size_t dstep);
// cv::remap(src, dst, mapx, mapy, INTER_LINEAR) with CV_32FC1 maps, BORDER_CONSTANT(0), 8-bit single channel.
void remap_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, int dw,
                     int dh, uint8_t* dst, size_t dstep);

// cv::undistortPoints(src, dst, K, distCoeffs, noArray(), K) for float points (OpenCV undistort.dispatch.cpp
// cvUndistortPointsInternal: 5 fixed-point iterations in double, then re-projection with K).  dist = up to 14
// coefficients in OpenCV order (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 ...); xy / out = n x 2 floats.
void undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, const double* dist, int ndist,
                      float* out);

}  // namespace cvp
