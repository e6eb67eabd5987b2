// tests/native/host_adapter_check.cpp -- compile AND run check of the header-only C++ adapter
// (orb-slam3_byzyh_b200/host/*.h) with ORB-SLAM3's unchanged signatures, against the cv::Mat /
// cv::KeyPoint layout stub oracle/cvshim (OpenCV C++ is not in this image).  Run on the GPU box
// by tests/test_gpu_host_adapter.py; compile-only on CPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher_b200.h"

// argv: rows cols nfeatures lap0 lap1 in.raw out_prefix
int main(int argc, char** argv) {
    if (argc < 8) return 2;
    const int rows = atoi(argv[1]), cols = atoi(argv[2]), nf = atoi(argv[3]);
    std::vector<int> lap = {atoi(argv[4]), atoi(argv[5])};
    std::vector<unsigned char> img((size_t)rows * cols);
    FILE* f = fopen(argv[6], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 3;
    fclose(f);
    ORB_SLAM3::ORBextractor ex(nf, 1.2f, 8, 20, 7);
    cv::Mat im(rows, cols, CV_8UC1, img.data(), (size_t)cols), mask, desc;
    std::vector<cv::KeyPoint> kps;
    const int mono = ex(im, mask, kps, desc, lap);
    std::string p(argv[7]);
    f = fopen((p + ".kps").c_str(), "wb");
    fwrite(kps.data(), sizeof(cv::KeyPoint), kps.size(), f);
    fclose(f);
    f = fopen((p + ".desc").c_str(), "wb");
    for (int i = 0; i < desc.rows; i++) fwrite(desc.ptr(i), 1, 32, f);
    fclose(f);
    // mvImagePyramid: ROI with readable border, like the reference's
    const cv::Mat& l1 = ex.mvImagePyramid[1];
    f = fopen((p + ".pyr1").c_str(), "wb");
    for (int y = -19; y < l1.rows + 19; y++) fwrite(l1.data + (long)y * (long)(size_t)l1.step - 19, 1, l1.cols + 38, f);
    fclose(f);
    int self = desc.rows ? ORB_SLAM3::b200::DescriptorDistance(desc.row(0), desc.row(0)) : 0;
    printf("%d %zu %d %d %d %.3f\n", mono, kps.size(), ex.GetLevels(), l1.cols, self, ex.GetScaleFactor());
    return 0;
}
