// tests/native/host_adapter_check.cpp -- compile AND run check of the header-only C++ adapter
// (orb-slam3_byzyh_b200/host/*.h) with ORB-SLAM3's unchanged signatures, against the cv::Mat /
// cv::KeyPoint layout stub oracle/cvshim (OpenCV C++ is not in this image).  Run on the GPU box
// by tests/test_gpu_host_adapter.py; compile-only on CPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
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
    // matcher adapter: the frame's own keypoints re-projected onto themselves (1 px off), as map points
    {
        using namespace ORB_SLAM3::b200;
        const int n = (int)kps.size();
        ProjPoints P;
        for (int i = 0; i < n; i++)
            P.push(kps[i].pt.x + 1.0f, kps[i].pt.y - 1.0f, kps[i].pt.x - 4.0f, 6.0f * ex.GetScaleFactors()[kps[i].octave],
                   kps[i].octave - 1, kps[i].octave, kps[i].angle, (i % 17) != 0, true, desc.row(i));
        std::vector<float> uright(n);
        for (int i = 0; i < n; i++) uright[i] = (i % 3) ? kps[i].pt.x - 5.0f : -1.0f;
        const float wInv = 64.0f / cols, hInv = 48.0f / rows;
        KeyFrameView kv{&kps, &uright, desc.ptr(), 0.f, 0.f, (float)cols, (float)rows, wInv, hInv};
        std::vector<int32_t> fuse, sim3, assigned(n, -2);
        std::vector<float> inv = ex.GetInverseScaleSigmaSquares();
        const int nf1 = FuseSearch(kv, P, inv, fuse);
        const int ns = SearchBySim3(kv, kv, P, P, sim3);
        std::vector<uint8_t> claimed(n, 0);
        const int np = SearchByProjection(kps, uright, desc, 0.f, 0.f, (float)cols, (float)rows, wInv, hInv, P,
                                          ORBFE_SEARCH_MAPPOINTS, 100, 0.8f, false, claimed, assigned);
        std::vector<int32_t> asim(n, -2);
        for (int i = 0; i < n; i++) claimed[i] = (i % 5) == 0;
        const int nk = SearchByProjectionSim3(kv, P, 1.0f, claimed, asim);
        f = fopen((p + ".match").c_str(), "wb");
        fwrite(fuse.data(), 4, n, f); fwrite(sim3.data(), 4, n, f); fwrite(assigned.data(), 4, n, f); fwrite(asim.data(), 4, n, f);
        fclose(f);
        fprintf(stderr, "fuse %d sim3 %d proj %d kfsim3 %d\n", nf1, ns, np, nk);
    }
    // vocabulary adapter: argv[8] = vocabulary text file (optional); BoW of the frame and a KeyFrame-Frame search
    // of the frame against itself
    if (argc > 8) {
        using namespace ORB_SLAM3::b200;
        ORBVocabulary voc;
        if (!voc.loadFromTextFile(argv[8])) return 4;
        std::map<unsigned, double> bow;
        std::map<unsigned, std::vector<unsigned>> fv;
        voc.transform(desc, bow, fv, 2);
        const int n = (int)kps.size();
        std::vector<float> ang(n);
        for (int i = 0; i < n; i++) ang[i] = kps[i].angle;
        std::vector<uint8_t> valid(n), none;
        for (int i = 0; i < n; i++) valid[i] = (i % 7) != 0;
        std::vector<int32_t> mA, mR;
        const int nb = SearchByBoW(desc, ang, valid, fv, desc, ang, none, fv, 0.9f, true, false, -1, mA, mR);
        f = fopen((p + ".bow").c_str(), "wb");
        for (auto& e : bow) { fwrite(&e.first, 4, 1, f); fwrite(&e.second, 8, 1, f); }
        fclose(f);
        f = fopen((p + ".bowmatch").c_str(), "wb");
        fwrite(mA.data(), 4, n, f);
        fclose(f);
        // SearchForTriangulation of the frame against itself (coarse: no epipolar gate; every other slot "has a point")
        std::vector<uint8_t> hasMp(n);
        for (int i = 0; i < n; i++) hasMp[i] = i & 1;
        std::vector<float> noR, sf = ex.GetScaleFactors(), s2 = ex.GetScaleSigmaSquares();
        const float F12[9] = {0, 0, 0, 0, 0, -1, 0, 1, 0}, ep[2] = {-1000.f, -1000.f};
        std::vector<std::pair<size_t, size_t> > pairs;
        const int nt = SearchForTriangulation(kps, desc, noR, hasMp, fv, kps, desc, noR, hasMp, fv, F12, ep, sf, s2, false, true, true,
                                              pairs);
        f = fopen((p + ".tri").c_str(), "wb");
        for (auto& pr : pairs) { int32_t ab[2] = {(int32_t)pr.first, (int32_t)pr.second}; fwrite(ab, 4, 2, f); }
        fclose(f);
        fprintf(stderr, "bow words %zu nodes %zu matches %d tri %d\n", bow.size(), fv.size(), nb, nt);
        {   // fisheye triangulation of the frame's keypoints against themselves shifted along a 10 cm baseline: the call
            // must go through (every pair is either accepted with a positive depth or carries a rejection code)
            const float P[8] = {190.97f, 190.97f, 254.93f, 256.89f, 0.0034f, 0.0007f, -0.0020f, 0.0002f};
            const float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0.1f, 0, 0};
            std::vector<std::pair<int, int> > fpairs;
            for (int i = 0; i + 1 < (int)kps.size() && i < 64; i++) fpairs.push_back(std::make_pair(i, i + 1));
            std::vector<float> depth, p3D;
            TriangulateFisheyeMatches(P, 1e-6f, P, 1e-6f, R, t, kps, kps, fpairs, s2, depth, p3D);
            for (size_t i = 0; i < depth.size(); i++)
                if (!(depth[i] > 0.f || (depth[i] <= -1.f && depth[i] >= -5.f))) return 7;
        }
    }
    int self = desc.rows ? ORB_SLAM3::b200::DescriptorDistance(desc.row(0), desc.row(0)) : 0;
    printf("%d %zu %d %d %d %.3f\n", mono, kps.size(), ex.GetLevels(), l1.cols, self, ex.GetScaleFactor());
    return 0;
}
