// oracle/cvshim/cvshim.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the slice of the OpenCV C++ API that the reference's
// src/ORBextractor.cc touches, so that file can be compiled *verbatim* (from where it lies
// under /root/reference, never copied) into oracle/_ref/libref_orbextractor.so.  OpenCV C++
// headers/libraries are not in this image; the five image primitives are forwarded to the
// restatements in oracle/cvprims.* (pinned bit-exact against cv2 4.13.0 by the tests).
// The same shim lets the product's C++ adapter (orb-slam3_byzyh_b200/host) be compile- and
// run-checked with ORB-SLAM3's unchanged signatures.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdlib>
#include <iostream>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../cvprims.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

inline int cvRound(float v) { return cvp::cvRound(v); }
inline int cvRound(double v) { return cvp::cvRound(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return cvp::cvFloor(v); }
inline int cvCeil(double v) { return cvp::cvCeil(v); }

namespace cv {

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U>
    Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
template <typename T>
inline Point_<T>& operator*=(Point_<T>& a, float b) {
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}
template <typename T>
inline Point_<T>& operator*=(Point_<T>& a, double b) {
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};
struct Range {
    int start, end;
    Range(int s, int e) : start(s), end(e) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatStep {
    size_t v;
    MatStep() : v(0) {}
    MatStep(size_t s) : v(s) {}
    operator size_t() const { return v; }
};

// Single-channel 8-bit matrix with OpenCV's shared-buffer / ROI semantics.
class Mat {
   public:
    int rows, cols;
    uchar* data;
    MatStep step;
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() : rows(0), cols(0), data(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), data(nullptr) { create(sz.height, sz.width, type); }
    // wrap user memory (no ownership), like cv::Mat(rows, cols, type, ptr, step)
    Mat(int r, int c, int /*type*/, void* ptr, size_t s = 0)
        : rows(r), cols(c), data((uchar*)ptr), step(s ? s : (size_t)c) {}

    void create(int r, int c, int /*type*/) {
        if (data && r == rows && c == cols) return;
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c);
        data = buf->data();
        rows = r; cols = c; step = (size_t)c;
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        if (m.data) memset(m.data, 0, (size_t)r * c);
        return m;
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    int channels() const { return 1; }
    Size size() const { return Size(cols, rows); }
    size_t step1() const { return step; }
    bool isContinuous() const { return (size_t)cols == (size_t)step || rows == 1; }

    Mat operator()(const Rect& r) const {
        Mat m;
        m.buf = buf; m.rows = r.height; m.cols = r.width; m.step = step;
        m.data = data + (size_t)r.y * step + r.x;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat clone() const {
        Mat m(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    void copyTo(const Mat& dst) const {
        assert(dst.rows == rows && dst.cols == cols);
        for (int y = 0; y < rows; y++) memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, cols);
    }
    template <typename T>
    T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T>
    const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x * sizeof(T)); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T>
    T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T>
    const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
};

class _InputArray {
   public:
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }

   private:
    const Mat* m_;
};
class _OutputArray {
   public:
    _OutputArray(Mat& m) : m_(&m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
    Mat getMat() const { return *m_; }
    Mat& ref() const { return *m_; }

   private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interp = INTER_LINEAR) {
    assert(interp == INTER_LINEAR);
    dst.create(dsize.height, dsize.width, CV_8UC1);
    cvp::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    cvp::copy_make_border_reflect101(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top,
                                     bottom, left, right);
}
inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nms = true) {
    std::vector<cvp::FastKP> out;
    cvp::fast9_16(img.data, img.cols, img.rows, img.step, threshold, nms, out);
    kps.clear();
    for (const auto& k : out) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1, (float)k.score));
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy = 0,
                         int borderType = BORDER_DEFAULT) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    Mat tmp = src.clone();  // in-place call at reference src/ORBextractor.cc:1632
    dst.create(src.rows, src.cols, CV_8UC1);
    cvp::gaussian_blur_7x7_s2(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step);
}
inline float fastAtan2(float y, float x) { return cvp::fast_atan2(y, x); }

// cv::norm(a, b, NORM_L1) on 8-bit single-channel views (reference src/Frame.cc:1284): sum |a-b|,
// an exact integer in double.
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
inline double norm(const Mat& a, const Mat& b, int normType) {
    assert(normType == NORM_L1 && a.rows == b.rows && a.cols == b.cols);
    (void)normType;
    long long s = 0;
    for (int y = 0; y < a.rows; y++) {
        const uchar *pa = a.ptr(y), *pb = b.ptr(y);
        for (int x = 0; x < a.cols; x++) s += pa[x] > pb[x] ? pa[x] - pb[x] : pb[x] - pa[x];
    }
    return (double)s;
}

// cv::FileStorage / cv::FileNode: DBoW2's TemplatedVocabulary.h has YAML save/load members (virtual, so they
// must compile) next to the text-file loader ORB-SLAM3 actually uses (System.cc:105 loadFromTextFile); here
// they only have to compile and are never called.
struct FileNode {
    enum { NONE = 0, SEQ = 5, MAP = 6 };
    FileNode operator[](const char*) const { abort(); }
    FileNode operator[](const std::string&) const { abort(); }
    FileNode operator[](int) const { abort(); }
    operator int() const { abort(); }
    operator double() const { abort(); }
    operator std::string() const { abort(); }
    size_t size() const { abort(); }
    int type() const { abort(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { abort(); }
    FileNode operator[](const std::string&) const { abort(); }
    template <class T> FileStorage& operator<<(const T&) { abort(); }
};

// Only named by the reference's dead ComputeKeyPointsOld (call commented out at
// src/ORBextractor.cc:1580); never executed.
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>&, int) { abort(); }
};

}  // namespace cv
