// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the reference's own ORB_SLAM3::ORBextractor (src/ORBextractor.cc,
// compiled verbatim against oracle/cvshim by oracle/ref_build.sh into oracle/_ref/).  Used to
// pin the restatement in orb_oracle.cpp and, on the GPU box, as the CPU baseline.
#include <cstring>
#include <vector>

#include "ORBextractor.h"

namespace {
class RefExtractor : public ORB_SLAM3::ORBextractor {
   public:
    using ORB_SLAM3::ORBextractor::ORBextractor;
    std::vector<cv::KeyPoint> octree(const std::vector<cv::KeyPoint>& v, int minX, int maxX, int minY,
                                     int maxY, int N) {
        return DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
    }
};
}  // namespace

extern "C" {
void* ref_extractor_create(int nf, float sf, int nl, int ini, int mn) {
    return new RefExtractor(nf, sf, nl, ini, mn);
}
void ref_extractor_destroy(void* h) { delete (RefExtractor*)h; }

// kps: cv::KeyPoint[cap] (28 B each); desc: cap x 32.  Returns monoIndex / -1.
int ref_extract(void* h, const unsigned char* img, int rows, int cols, int step, int lap0, int lap1,
                void* kps, unsigned char* desc, int cap, int* n_out) {
    RefExtractor* e = (RefExtractor*)h;
    cv::Mat image = (rows > 0 && cols > 0) ? cv::Mat(rows, cols, CV_8UC1, (void*)img, (size_t)step) : cv::Mat();
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    std::vector<int> lap = {lap0, lap1};
    int mono = (*e)(image, cv::Mat(), k, d, lap);
    *n_out = (int)k.size();
    if (mono < 0) { *n_out = 0; return mono; }
    int n = (int)k.size() < cap ? (int)k.size() : cap;
    if (n) {
        memcpy(kps, k.data(), sizeof(cv::KeyPoint) * n);
        for (int i = 0; i < n; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    }
    return mono;
}
// Bordered pyramid level of the last call: (h+38) x (w+38) bytes, row pitch w+38.
void ref_level_dims(void* h, int lvl, int* w, int* hh) {
    RefExtractor* e = (RefExtractor*)h;
    *w = e->mvImagePyramid[lvl].cols; *hh = e->mvImagePyramid[lvl].rows;
}
void ref_level_padded(void* h, int lvl, unsigned char* out) {
    RefExtractor* e = (RefExtractor*)h;
    const cv::Mat& m = e->mvImagePyramid[lvl];
    const int W = m.cols + 38, H = m.rows + 38;
    const unsigned char* base = m.data - 19 * (size_t)m.step - 19;
    for (int y = 0; y < H; y++) memcpy(out + (size_t)y * W, base + (size_t)y * m.step, W);
}
// DistributeOctTree alone: xys = (x,y,score) in window coords -> kept (x,y,score), list order.
int ref_octree(void* h, const int* xys, int n, int minX, int maxX, int minY, int maxY, int N, int* out, int cap) {
    RefExtractor* e = (RefExtractor*)h;
    std::vector<cv::KeyPoint> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]);
    std::vector<cv::KeyPoint> r = e->octree(v, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) {
        out[3 * i] = (int)r[i].pt.x; out[3 * i + 1] = (int)r[i].pt.y; out[3 * i + 2] = (int)r[i].response;
    }
    return (int)r.size();
}
}
