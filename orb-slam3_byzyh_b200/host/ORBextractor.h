// host/ORBextractor.h -- drop-in replacement for ORB-SLAM3's include/ORBextractor.h
// (/root/reference/include/ORBextractor.h:46-112): the class keeps the reference's name, namespace,
// constructor, operator() and accessor signatures and its public mvImagePyramid member, so
// Frame.cc / Tracking.cc compile and link against it unchanged; every method forwards to the C ABI
// of libORBfe_b200.so (include/orbfe.h).  Header-only; needs OpenCV's core headers (or the layout
// stub oracle/cvshim used by the compile/run check in tests/test_host_adapter.py).
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include "orbfe.h"

namespace ORB_SLAM3 {

class ORBextractor {
   public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // ORBextractor.h:49-50.  The CUDA ordinal comes from ORBFE_DEVICE (default 0).
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
        : mbHostPyramid(true), h_(nullptr), nlevels(nlevels) {
        const char* d = getenv("ORBFE_DEVICE");
        if (orbfe_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, d ? atoi(d) : 0, &h_) != ORBFE_OK)
            throw std::runtime_error(std::string("ORBextractor (B200): ") + orbfe_last_error());
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
        mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        orbfe_scale_tables(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data());
        mvImagePyramid.resize(nlevels);
        capacity_ = orbfe_max_keypoints(h_);
    }
    ~ORBextractor() { orbfe_extractor_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // ORBextractor.h:57-59 / ORBextractor.cc:1557-1682.  Mask is ignored, as in the reference.
    int operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
        if (_image.empty()) return -1;  // :1561-1562
        cv::Mat image = _image.getMat();
        assert(image.type() == CV_8UC1);  // :1567
        static_assert(sizeof(cv::KeyPoint) == sizeof(OrbfeKeyPoint), "cv::KeyPoint layout");
        // frames wider than 8.5 : 1 start the quadtree from more roots than capacity_ allows for
        const int cap = std::max(capacity_, orbfe_max_keypoints_for(h_, image.rows, image.cols));
        std::vector<cv::KeyPoint> kps(cap);
        std::vector<unsigned char> desc((size_t)cap * 32);
        int n = 0;
        const int mono = orbfe_extract(h_, image.data, image.rows, image.cols, (size_t)image.step, vLappingArea[0],
                                       vLappingArea[1], reinterpret_cast<OrbfeKeyPoint*>(kps.data()), desc.data(),
                                       cap, &n);
        if (mono < 0) throw std::runtime_error(std::string("ORBextractor (B200): ") + orbfe_last_error());
        kps.resize(n);
        _keypoints.swap(kps);
        if (n == 0) {
            _descriptors.release();  // :1596-1597
        } else {
            _descriptors.create(n, 32, CV_8U);  // :1600
            cv::Mat d = _descriptors.getMat();
            for (int i = 0; i < n; i++) memcpy(d.ptr(i), &desc[(size_t)i * 32], 32);
        }
        if (mbHostPyramid) SyncPyramidToHost(rectRows_ ? rectRows_ : image.rows, rectCols_ ? rectCols_ : image.cols);
        return mono;  // :1681
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return orbfe_get_scale_factor(h_); }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // ORBextractor.h:83.  Host mirror of the device pyramid: each level is an ROI inside a
    // (w+38) x (h+38) bordered buffer, exactly like the reference's (:1695-1716), so CPU code that
    // reads outside the ROI (Frame::ComputeStereoMatches windows) sees the same bytes.  Set
    // mbHostPyramid = false when stereo matching also runs on the device (ORBmatcher_b200.h): the
    // pyramid then never leaves HBM.
    std::vector<cv::Mat> mvImagePyramid;
    bool mbHostPyramid;

    void SyncPyramidToHost(int rows, int cols) {
        for (int l = 0; l < nlevels; l++) {
            int w = 0, hgt = 0;
            orbfe_level_size(h_, rows, cols, l, &w, &hgt);
            cv::Mat temp(hgt + 2 * ORBFE_EDGE, w + 2 * ORBFE_EDGE, CV_8UC1);
            if (orbfe_pyramid_level(h_, 0, l, 1, temp.data, (size_t)temp.step) != ORBFE_OK)
                throw std::runtime_error(std::string("ORBextractor (B200): ") + orbfe_last_error());
            mvImagePyramid[l] = temp(cv::Rect(ORBFE_EDGE, ORBFE_EDGE, w, hgt));
        }
    }

    // Stereo rectification fused into the extractor (System::TrackStereo, System.cc:286-293): with maps set
    // (M1 / M2 of Settings as CV_32FC1, rows x cols of the rectified image) operator() takes the RAW camera image and
    // mvImagePyramid[0] is the rectified image.  Null pointers switch it off.
    void SetRectification(const float* mapX, const float* mapY, int rows, int cols) {
        if (orbfe_extractor_set_rectification(h_, mapX, mapY, rows, cols) != ORBFE_OK)
            throw std::runtime_error(std::string("ORBextractor (B200): ") + orbfe_last_error());
        rectRows_ = mapX ? rows : 0; rectCols_ = mapX ? cols : 0;
    }

    OrbfeExtractor* handle() const { return h_; }  // for ORBmatcher_b200.h

   protected:
    OrbfeExtractor* h_;
    int nlevels;
    int capacity_;
    int rectRows_ = 0, rectCols_ = 0;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM3

#endif
