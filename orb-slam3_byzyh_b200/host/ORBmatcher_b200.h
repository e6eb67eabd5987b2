// host/ORBmatcher_b200.h -- the Hamming-matching entry points of ORB-SLAM3 on the C ABI.
//
// ORBmatcher::SearchByProjection and Frame::ComputeStereoMatches are member functions that walk
// Frame / MapPoint / KeyFrame object graphs (Eigen, Sophus, mutexes); the boundary
// keeps that walking on the host and hands plain arrays to the device.  This header provides the
// array-level calls the three-line patches in INTEGRATION.md insert into src/ORBmatcher.cc and
// src/Frame.cc; it depends only on <vector>, OpenCV core types and include/orbfe.h.
#ifndef ORBMATCHER_B200_H
#define ORBMATCHER_B200_H

#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <opencv2/opencv.hpp>

#include "ORBextractor.h"
#include "orbfe.h"

namespace ORB_SLAM3 {
namespace b200 {

inline int device() {
    const char* d = getenv("ORBFE_DEVICE");
    return d ? atoi(d) : 0;
}

// static int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b)   ORBmatcher.h:43
inline int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    int32_t out = 0;
    if (orbfe_descriptor_distance(a.ptr(), b.ptr(), 1, &out, device()) != ORBFE_OK)
        throw std::runtime_error(std::string("DescriptorDistance (B200): ") + orbfe_last_error());
    return out;
}

// One projected candidate as the reference loops see it; the caller evaluates the early
// `continue`s of its loop (bad point, not in view, depth, out of image) into `valid`.
struct ProjPoints {
    std::vector<float> u, v, ur, radius, angle;
    std::vector<int32_t> minLevel, maxLevel;
    std::vector<uint8_t> valid, blocks, desc;  // desc: 32 bytes per point (MapPoint::GetDescriptor())
    void push(float u_, float v_, float ur_, float r_, int minL, int maxL, float ang, bool ok, bool blk,
              const cv::Mat& d) {
        u.push_back(u_); v.push_back(v_); ur.push_back(ur_); radius.push_back(r_); angle.push_back(ang);
        minLevel.push_back(minL); maxLevel.push_back(maxL); valid.push_back(ok); blocks.push_back(blk);
        const unsigned char* p = d.ptr();
        desc.insert(desc.end(), p, p + 32);
    }
    size_t size() const { return u.size(); }
};

// keys = F.mvKeysUn, uright = F.mvuRight (empty for monocular), desc = F.mDescriptors (N x 32,
// continuous), bounds = F.mnMinX/MinY/MaxX/MaxY, inv = F.mfGridElementWidthInv/HeightInv.
// claimed[i] != 0 <=> F.mvpMapPoints[i] && Observations() > 0 (modes 0/1) or non-null (mode 2).
// assigned[i] receives the index (into pts) of the point the reference would store in
// F.mvpMapPoints[i]; -1 = cleared by the rotation check; untouched entries keep their value.
inline int SearchByProjection(const std::vector<cv::KeyPoint>& keys, const std::vector<float>& uright,
                              const cv::Mat& desc, float minX, float minY, float maxX, float maxY, float wInv,
                              float hInv, const ProjPoints& pts, int mode, int thAccept, float nnratio,
                              bool checkOrientation, const std::vector<uint8_t>& claimed,
                              std::vector<int32_t>& assigned) {
    OrbfeFrameView fv;
    fv.n = (int32_t)keys.size();
    fv.keys = reinterpret_cast<const OrbfeKeyPoint*>(keys.data());
    fv.uright = uright.empty() ? nullptr : uright.data();
    fv.desc = desc.ptr();
    fv.min_x = minX; fv.min_y = minY; fv.max_x = maxX; fv.max_y = maxY;
    fv.grid_w_inv = wInv; fv.grid_h_inv = hInv;
    OrbfeProjPoints pp;
    pp.m = (int32_t)pts.size();
    pp.u = pts.u.data(); pp.v = pts.v.data(); pp.ur = pts.ur.data(); pp.radius = pts.radius.data();
    pp.min_level = pts.minLevel.data(); pp.max_level = pts.maxLevel.data(); pp.angle = pts.angle.data();
    pp.valid = pts.valid.data(); pp.blocks = pts.blocks.data(); pp.desc = pts.desc.data();
    OrbfeSearchParams prm = {mode, thAccept, nnratio, checkOrientation ? 1 : 0};
    const int n = orbfe_search_by_projection(&fv, &pp, &prm, claimed.data(), assigned.data(), nullptr, nullptr, device());
    if (n < 0) throw std::runtime_error(std::string("SearchByProjection (B200): ") + orbfe_last_error());
    return n;
}

// The same two overloads on a fisheye stereo frame (F.Nleft != -1): keysL = F.mvKeys, keysR = F.mvKeysRight,
// desc = F.mDescriptors (rows [0,Nleft) left, [Nleft,N) right), l2r / r2l = F.mvLeftToRightMatch /
// F.mvRightToLeftMatch; ptsL carries the left projections and the shared angle / blocks / desc, ptsR the
// right-camera projections (ORBmatcher.cc:171-237, 2090-2155).  claimed / assigned are indexed like
// F.mvpMapPoints.  mode = ORBFE_SEARCH_MAPPOINTS or ORBFE_SEARCH_LASTFRAME.
inline int SearchByProjectionFisheye(const std::vector<cv::KeyPoint>& keysL, const std::vector<cv::KeyPoint>& keysR,
                                     const cv::Mat& desc, float minX, float minY, float maxX, float maxY, float wInv,
                                     float hInv, const std::vector<int>& l2r, const std::vector<int>& r2l,
                                     const ProjPoints& ptsL, const ProjPoints& ptsR, int mode, int thAccept,
                                     float nnratio, bool checkOrientation, const std::vector<uint8_t>& claimed,
                                     std::vector<int32_t>& assigned) {
    OrbfeFrameView fl, fr;
    fl.n = (int32_t)keysL.size(); fl.keys = reinterpret_cast<const OrbfeKeyPoint*>(keysL.data()); fl.uright = nullptr;
    fl.desc = desc.ptr();
    fr.n = (int32_t)keysR.size(); fr.keys = reinterpret_cast<const OrbfeKeyPoint*>(keysR.data()); fr.uright = nullptr;
    fr.desc = desc.ptr() + (size_t)32 * keysL.size();
    fl.min_x = fr.min_x = minX; fl.min_y = fr.min_y = minY; fl.max_x = fr.max_x = maxX; fl.max_y = fr.max_y = maxY;
    fl.grid_w_inv = fr.grid_w_inv = wInv; fl.grid_h_inv = fr.grid_h_inv = hInv;
    auto view = [](const ProjPoints& p, const ProjPoints& shared) {
        OrbfeProjPoints pp;
        pp.m = (int32_t)p.size();
        pp.u = p.u.data(); pp.v = p.v.data(); pp.ur = nullptr; pp.radius = p.radius.data();
        pp.min_level = p.minLevel.data(); pp.max_level = p.maxLevel.data(); pp.angle = shared.angle.data();
        pp.valid = p.valid.data(); pp.blocks = shared.blocks.data(); pp.desc = shared.desc.data();
        return pp;
    };
    const OrbfeProjPoints pl = view(ptsL, ptsL), pr = view(ptsR, ptsL);
    OrbfeSearchParams prm = {mode, thAccept, nnratio, checkOrientation ? 1 : 0};
    const int n = orbfe_search_by_projection_fisheye(&fl, &fr, l2r.data(), r2l.data(), &pl, &pr, &prm, claimed.data(),
                                                     assigned.data(), nullptr, nullptr, device());
    if (n < 0) throw std::runtime_error(std::string("SearchByProjection fisheye (B200): ") + orbfe_last_error());
    return n;
}

// A keyframe as the keyframe-side searches see it: pKF->mvKeysUn (mvKeys / mvKeysRight for a fisheye
// keyframe), pKF->mvuRight, pKF->mDescriptors and the KeyFrame's grid bounds (mnMinX.. are ints there).
struct KeyFrameView {
    const std::vector<cv::KeyPoint>* keys;
    const std::vector<float>* uright;   // may be null
    const uint8_t* desc;
    float minX, minY, maxX, maxY, wInv, hInv;
    OrbfeFrameView view() const {
        OrbfeFrameView fv;
        fv.n = (int32_t)keys->size();
        fv.keys = reinterpret_cast<const OrbfeKeyPoint*>(keys->data());
        fv.uright = (uright && !uright->empty()) ? uright->data() : nullptr;
        fv.desc = desc;
        fv.min_x = minX; fv.min_y = minY; fv.max_x = maxX; fv.max_y = maxY;
        fv.grid_w_inv = wInv; fv.grid_h_inv = hInv;
        return fv;
    }
};

inline OrbfeProjPoints view(const ProjPoints& p) {
    OrbfeProjPoints pp;
    pp.m = (int32_t)p.size();
    pp.u = p.u.data(); pp.v = p.v.data(); pp.ur = p.ur.data(); pp.radius = p.radius.data();
    pp.min_level = p.minLevel.data(); pp.max_level = p.maxLevel.data(); pp.angle = p.angle.data();
    pp.valid = p.valid.data(); pp.blocks = p.blocks.data(); pp.desc = p.desc.data();
    return pp;
}

// Inner loop of int ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight)  (ORBmatcher.cc:1326-1534,
// reprojection gate :1436-1461 -> pass pKF->mvInvLevelSigma2) and of Fuse(KeyFrame*, Sim3f&, ...) (:1536-1688,
// no gate -> pass an empty vector).  pts: one entry per candidate map point (level window [nPredictedLevel-1,
// nPredictedLevel], ur = uv(0) - bf*invz, valid = passed :1360-1407).  bestIdx[i] = keypoint the point fuses
// with (bestDist <= TH_LOW) or -1; the caller then runs :1480-1501 / :1645-1661 on its MapPoint graph:
//     for i in order, bestIdx[i] >= 0:  pMPinKF = pKF->GetMapPoint(bestIdx[i]);
//         pMPinKF ? (Replace / vpReplacePoint[i] = pMPinKF) : (AddObservation + AddMapPoint);  nFused++
inline int FuseSearch(const KeyFrameView& kf, const ProjPoints& pts, const std::vector<float>& invLevelSigma2,
                      std::vector<int32_t>& bestIdx, int thLow = 50) {
    bestIdx.assign(pts.size(), -1);
    if (pts.size() == 0) return 0;
    const OrbfeFrameView fv = kf.view();
    const OrbfeProjPoints pp = view(pts);
    OrbfeWindowParams prm = {thLow, invLevelSigma2.empty() ? ORBFE_GATE_NONE : ORBFE_GATE_FUSE,
                             invLevelSigma2.empty() ? nullptr : invLevelSigma2.data(), (int32_t)invLevelSigma2.size()};
    const int n = orbfe_search_window(&fv, &pp, &prm, bestIdx.data(), nullptr, device());
    if (n < 0) throw std::runtime_error(std::string("Fuse (B200): ") + orbfe_last_error());
    return n;
}

// int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const Sim3f& S12, th)
// (ORBmatcher.cc:1690-1940).  pts12[i1]: KF1's map point i1 through S21 into KF2 (valid = :1731-1765 passed, i.e.
// not in vbAlreadyMatched1), pts21 the reverse.  match12[i1] = idx2 agreed by both directions or -1; the caller
// stores vpMatches12[i1] = vpMapPoints2[match12[i1]].  Returns nFound.
inline int SearchBySim3(const KeyFrameView& kf1, const KeyFrameView& kf2, const ProjPoints& pts12,
                        const ProjPoints& pts21, std::vector<int32_t>& match12, int thHigh = 100) {
    match12.assign(kf1.keys->size(), -1);
    const OrbfeFrameView f1 = kf1.view(), f2 = kf2.view();
    const OrbfeProjPoints p12 = view(pts12), p21 = view(pts21);
    const int n = orbfe_search_by_sim3(&f1, &f2, &p12, &p21, thHigh, match12.data(), device());
    if (n < 0) throw std::runtime_error(std::string("SearchBySim3 (B200): ") + orbfe_last_error());
    return n;
}

// int ORBmatcher::SearchByProjection(KeyFrame* pKF, Sim3f& Scw, const vector<MapPoint*>& vpPoints,
// vector<MapPoint*>& vpMatched, int th, float ratioHamming)  (ORBmatcher.cc:496-610; the overload with
// vpPointsKFs :612-733 runs the same loop).  matched[i] != 0 <=> vpMatched[i] != NULL on entry; assigned[i]
// (size pKF->N, initialised to -2) = index into vpPoints of the point now in vpMatched[i].  Returns nmatches.
inline int SearchByProjectionSim3(const KeyFrameView& kf, const ProjPoints& pts, float ratioHamming,
                                  const std::vector<uint8_t>& matched, std::vector<int32_t>& assigned, int thLow = 50) {
    const OrbfeFrameView fv = kf.view();
    const OrbfeProjPoints pp = view(pts);
    // bestDist <= TH_LOW*ratioHamming (int vs float, :589)  <=>  bestDist <= floor(TH_LOW*ratioHamming)
    OrbfeSearchParams prm = {ORBFE_SEARCH_KEYFRAME, (int32_t)std::floor((float)thLow * ratioHamming), 1.0f, 0};
    const int n = orbfe_search_by_projection(&fv, &pp, &prm, matched.data(), assigned.data(), nullptr, nullptr, device());
    if (n < 0) throw std::runtime_error(std::string("SearchByProjection Sim3 (B200): ") + orbfe_last_error());
    return n;
}

// ---- Bag of words --------------------------------------------------------------------------------------------
// ORBVocabulary (include/ORBVocabulary.h:30-31) with the tree in HBM.  transform() has the signature of
// DBoW2::TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) and is generic over the two
// map types (DBoW2::BowVector = std::map<WordId, WordValue>, DBoW2::FeatureVector = std::map<NodeId, vector<unsigned>>),
// so this header does not need DBoW2's own headers; Frame::ComputeBoW (Frame.cc:984-998) and KeyFrame::ComputeBoW
// (KeyFrame.cc:101-111) call it with mDescriptors instead of the vector<cv::Mat> copy.
class ORBVocabulary {
   public:
    ORBVocabulary() {}
    ~ORBVocabulary() { orbfe_vocabulary_destroy(h_); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;

    // ORBvoc.txt (TemplatedVocabulary::loadFromTextFile, :1338-1424; called at System.cc:105)
    bool loadFromTextFile(const std::string& filename) {
        FILE* f = fopen(filename.c_str(), "r");
        if (!f) return false;
        int n1 = 0, n2 = 0;
        if (fscanf(f, "%d %d %d %d", &k_, &L_, &n1, &n2) != 4 || k_ < 0 || k_ > 20 || L_ < 1 || L_ > 10 || n1 < 0 || n1 > 5 ||
            n2 < 0 || n2 > 3) { fclose(f); return false; }
        scoring_ = n1; weighting_ = n2;
        std::vector<int32_t> parent(1, 0);
        std::vector<uint8_t> desc(32, 0);
        std::vector<double> weight(1, 0.0);
        for (;;) {
            int pid, leaf;
            if (fscanf(f, "%d %d", &pid, &leaf) != 2) break;     // a trailing empty line ends the file here
            parent.push_back(pid);
            for (int i = 0; i < 32; i++) { int b = 0; if (fscanf(f, "%d", &b) != 1) b = 0; desc.push_back((uint8_t)b); }
            double w = 0;
            if (fscanf(f, "%lf", &w) != 1) w = 0;
            weight.push_back(w);
        }
        fclose(f);
        return create(k_, L_, (int)parent.size(), parent.data(), desc.data(), weight.data(), scoring_, weighting_);
    }
    bool create(int k, int L, int nNodes, const int32_t* parent, const uint8_t* desc, const double* weight, int scoring = 0,
                int weighting = 0) {
        orbfe_vocabulary_destroy(h_);
        h_ = nullptr;
        k_ = k; L_ = L; scoring_ = scoring; weighting_ = weighting;
        return orbfe_vocabulary_create(k, L, nNodes, parent, desc, weight, device(), &h_) == ORBFE_OK;
    }
    bool empty() const { return h_ == nullptr; }
    OrbfeVocabulary* handle() const { return h_; }

    // descriptors: N x 32 continuous (Frame::mDescriptors)
    template <class BowVector, class FeatureVector>
    void transform(const cv::Mat& descriptors, BowVector& v, FeatureVector& fv, int levelsup) const {
        v.clear();
        fv.clear();
        const int n = descriptors.rows;
        if (empty() || n == 0) return;
        std::vector<int32_t> word(n), node(n);
        std::vector<double> w(n);
        if (orbfe_bow_transform(h_, descriptors.ptr(), n, levelsup, word.data(), w.data(), node.data()) != ORBFE_OK)
            throw std::runtime_error(std::string("ORBVocabulary::transform (B200): ") + orbfe_last_error());
        const bool tf = weighting_ == 0 || weighting_ == 1;           // TF_IDF, TF  (DBoW2/BowVector.h:39-45)
        const bool must = scoring_ != 5;                              // every scoring but DOT_PRODUCT normalises
        for (int i = 0; i < n; i++) {
            if (!(w[i] > 0)) continue;                                // stopped word
            if (tf) v[word[i]] += w[i];                               // BowVector::addWeight
            else v.insert(typename BowVector::value_type(word[i], w[i]));   // addIfNotExist
            fv[node[i]].push_back((unsigned int)i);                   // FeatureVector::addFeature
        }
        if (tf && !v.empty() && !must) {
            const double nd = (double)v.size();
            for (auto& e : v) e.second /= nd;
        }
        if (must) {                                                   // BowVector::normalize (BowVector.cpp:62-84)
            double norm = 0.0;
            if (scoring_ == 1) { for (auto& e : v) norm += e.second * e.second; norm = std::sqrt(norm); }
            else for (auto& e : v) norm += std::fabs(e.second);
            if (norm > 0.0) for (auto& e : v) e.second /= norm;
        }
    }

   private:
    OrbfeVocabulary* h_ = nullptr;
    int k_ = 0, L_ = 0, scoring_ = 0, weighting_ = 0;
};

// Flattens a DBoW2::FeatureVector for the C ABI (keeps the storage alive).
struct FlatFeatureVector {
    std::vector<int32_t> node, start, feat;
    template <class FeatureVector>
    explicit FlatFeatureVector(const FeatureVector& fv) {
        start.push_back(0);
        for (const auto& e : fv) {
            node.push_back((int32_t)e.first);
            for (unsigned int i : e.second) feat.push_back((int32_t)i);
            start.push_back((int32_t)feat.size());
        }
    }
    OrbfeFeatureVector view() const { return OrbfeFeatureVector{(int32_t)node.size(), node.data(), start.data(), feat.data()}; }
};

// int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)  (ORBmatcher.cc:260-494)
//   a: pKF->mDescriptors / keypoint angles / validA[i] = (vpMapPointsKF[i] && !isBad()) / pKF->mFeatVec
//   b: F.mDescriptors / angles / F.mFeatVec, nLeft = F.Nleft
// and SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (:893-1044) with strict = true, validB = good points of
// pKF2.  matchA[ia] = matched index in b (left camera) or -1, matchAR[ia] = right-camera match (fisheye frames).
// Caller: vpMapPointMatches[matchA[ia]] = vpMapPointsKF[ia]   resp.   vpMatches12[ia] = vpMapPoints2[matchA[ia]].
template <class FeatureVector>
inline int SearchByBoW(const cv::Mat& descA, const std::vector<float>& angleA, const std::vector<uint8_t>& validA,
                       const FeatureVector& fvA, const cv::Mat& descB, const std::vector<float>& angleB,
                       const std::vector<uint8_t>& validB, const FeatureVector& fvB, float nnratio, bool checkOrientation,
                       bool strict, int nLeft, std::vector<int32_t>& matchA, std::vector<int32_t>& matchAR,
                       int thLow = 50) {
    const FlatFeatureVector fa(fvA), fb(fvB);
    OrbfeBowSide a = {descA.rows, descA.ptr(), angleA.data(), validA.empty() ? nullptr : validA.data(), fa.view()};
    OrbfeBowSide b = {descB.rows, descB.ptr(), angleB.data(), validB.empty() ? nullptr : validB.data(), fb.view()};
    matchA.assign(descA.rows, -1);
    matchAR.assign(descA.rows, -1);
    const int n = orbfe_search_by_bow(&a, &b, thLow, strict ? 1 : 0, nnratio, checkOrientation ? 1 : 0, nLeft, matchA.data(),
                                      matchAR.data(), device());
    if (n < 0) throw std::runtime_error(std::string("SearchByBoW (B200): ") + orbfe_last_error());
    return n;
}

// int ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse)  (ORBmatcher.cc:1046-1324).
// hasMapPoint*[i] = (pKF->GetMapPoint(i) != NULL); F12 = K1^-T * hat(t12) * R12 * K2^-1 (row major,
// the matrix Pinhole::epipolarConstrain builds, Pinhole.cpp:191-194) and epipole = pKF2->mpCamera->project(T2w * Cw)
// are computed once per keyframe pair by the caller (Eigen stays on the host).
// rig == nullptr: pinhole keyframes (mpCamera2 == NULL).  Two-camera keyframes (:1071-1095, :1160-1241): keys / desc =
// [mvKeys | mvKeysRight], uright empty, and rig = {NLeft of both keyframes, pKF1->mvLevelSigma2, and for ll, lr, rl, rr the
// KannalaBrandt8 parameters of the two cameras with R12 / t12 of Tll, Tlr, Trl, Trr}; F12 / epipole are then unused.
template <class FeatureVector>
inline int SearchForTriangulation(const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1, const std::vector<float>& uright1,
                                  const std::vector<uint8_t>& hasMapPoint1, const FeatureVector& fv1,
                                  const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2, const std::vector<float>& uright2,
                                  const std::vector<uint8_t>& hasMapPoint2, const FeatureVector& fv2, const float F12[9],
                                  const float epipole[2], const std::vector<float>& scaleFactors2,
                                  const std::vector<float>& levelSigma2_2, bool bOnlyStereo, bool bCoarse, bool checkOrientation,
                                  std::vector<std::pair<size_t, size_t> >& vMatchedPairs, int thLow = 50,
                                  const OrbfeTriRig* rig = nullptr) {
    const FlatFeatureVector f1(fv1), f2(fv2);
    OrbfeTriSide a = {(int32_t)keys1.size(), reinterpret_cast<const OrbfeKeyPoint*>(keys1.data()), desc1.ptr(),
                      uright1.empty() ? nullptr : uright1.data(), hasMapPoint1.data(), f1.view()};
    OrbfeTriSide b = {(int32_t)keys2.size(), reinterpret_cast<const OrbfeKeyPoint*>(keys2.data()), desc2.ptr(),
                      uright2.empty() ? nullptr : uright2.data(), hasMapPoint2.data(), f2.view()};
    OrbfeTriParams prm;
    for (int i = 0; i < 9; i++) prm.f12[i] = F12[i];
    prm.epipole[0] = epipole[0]; prm.epipole[1] = epipole[1];
    prm.scale_factors2 = scaleFactors2.data(); prm.level_sigma2_2 = levelSigma2_2.data(); prm.n_levels = (int32_t)scaleFactors2.size();
    prm.only_stereo = bOnlyStereo; prm.coarse = bCoarse; prm.check_orientation = checkOrientation; prm.th_low = thLow;
    prm.rig = rig;
    std::vector<int32_t> m12(keys1.size(), -1);
    const int n = orbfe_search_for_triangulation(&a, &b, &prm, m12.data(), device());
    if (n < 0) throw std::runtime_error(std::string("SearchForTriangulation (B200): ") + orbfe_last_error());
    vMatchedPairs.clear();
    vMatchedPairs.reserve(n);
    for (size_t i = 0; i < m12.size(); i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));   // :1316-1321
    return n;
}

// void MapPoint::ComputeDistinctiveDescriptors()  (MapPoint.cc:438-529) for many map points at once (LocalMapping calls it
// per new / fused point): descriptors of point p = rows start[p] .. start[p+1]-1 of `desc` in vDescriptors order;
// bestIdx[p] = the row to clone into mDescriptor (relative to start[p]), -1 when the point has no descriptor.
inline void ComputeDistinctiveDescriptors(const cv::Mat& desc, const std::vector<int32_t>& start, std::vector<int32_t>& bestIdx) {
    bestIdx.assign(start.empty() ? 0 : start.size() - 1, -1);
    if (bestIdx.empty()) return;
    if (orbfe_distinctive_descriptors(desc.ptr(), start.data(), (int)bestIdx.size(), bestIdx.data(), device()) != ORBFE_OK)
        throw std::runtime_error(std::string("ComputeDistinctiveDescriptors (B200): ") + orbfe_last_error());
}

// ---- Frame intake: the OpenCV calls made on an image before ORBextractor -------------------------------------------
// cv::cvtColor(im, im, cv::COLOR_{RGB,BGR,RGBA,BGRA}2GRAY)  (Tracking.cc:1563-1590, 1623-1636, 1702-1716)
inline void cvtColorToGray(const cv::Mat& src, cv::Mat& dst, int channels, bool rgbOrder) {
    cv::Mat out(src.rows, src.cols, CV_8UC1);
    if (orbfe_cvt_gray(src.data, src.rows, src.cols, (size_t)src.step, channels, rgbOrder ? 1 : 0, out.data, (size_t)out.step,
                       device()) != ORBFE_OK)
        throw std::runtime_error(std::string("cvtColor (B200): ") + orbfe_last_error());
    dst = out;
}
// cv::remap(imLeft, imLeftToFeed, M1l, M2l, cv::INTER_LINEAR)  (System.cc:292-293); M1 / M2 = the CV_32FC1 maps of
// Settings (x and y), passed as float pointers with rows x cols of the rectified image.
inline void remap(const cv::Mat& src, cv::Mat& dst, const float* mapX, const float* mapY, int rows, int cols) {
    cv::Mat out(rows, cols, CV_8UC1);
    if (orbfe_remap_linear(src.data, src.rows, src.cols, (size_t)src.step, mapX, mapY, rows, cols, out.data, (size_t)out.step,
                           device()) != ORBFE_OK)
        throw std::runtime_error(std::string("remap (B200): ") + orbfe_last_error());
    dst = out;
}
// void Frame::UndistortKeyPoints()  (Frame.cc:1003-1051): mvKeys -> mvKeysUn with K = mK (float entries), distCoef = mDistCoef
inline void UndistortKeyPoints(const std::vector<cv::KeyPoint>& keys, float fx, float fy, float cx, float cy,
                               const std::vector<float>& distCoef, std::vector<cv::KeyPoint>& keysUn) {
    keysUn.resize(keys.size());
    if (orbfe_undistort_keypoints(reinterpret_cast<const OrbfeKeyPoint*>(keys.data()), (int)keys.size(), fx, fy, cx, cy,
                                  distCoef.data(), (int)distCoef.size(), reinterpret_cast<OrbfeKeyPoint*>(keysUn.data()),
                                  device()) != ORBFE_OK)
        throw std::runtime_error(std::string("UndistortKeyPoints (B200): ") + orbfe_last_error());
}
// cv::resize(im, imToFeed, settings_->newImSize())  (System.cc:295-297)
inline void resize(const cv::Mat& src, cv::Mat& dst, cv::Size dsize) {
    cv::Mat out(dsize.height, dsize.width, CV_8UC1);
    if (orbfe_resize_linear(src.data, src.rows, src.cols, (size_t)src.step, dsize.height, dsize.width, out.data,
                            (size_t)out.step, device()) != ORBFE_OK)
        throw std::runtime_error(std::string("resize (B200): ") + orbfe_last_error());
    dst = out;
}

// void Frame::ComputeStereoMatches()   Frame.h:116, Frame.cc:1102-1358
inline void ComputeStereoMatches(ORBextractor* left, ORBextractor* right, const std::vector<cv::KeyPoint>& keysL,
                                 const cv::Mat& descL, const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR,
                                 float mbf, float mb, std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    mvuRight.assign(keysL.size(), -1.0f);
    mvDepth.assign(keysL.size(), -1.0f);
    if (keysL.empty()) return;
    if (orbfe_stereo_match(left->handle(), right->handle(), 0, reinterpret_cast<const OrbfeKeyPoint*>(keysL.data()),
                           descL.ptr(), (int)keysL.size(), reinterpret_cast<const OrbfeKeyPoint*>(keysR.data()),
                           descR.ptr(), (int)keysR.size(), mbf, mb, mvuRight.data(), mvDepth.data()) != ORBFE_OK)
        throw std::runtime_error(std::string("ComputeStereoMatches (B200): ") + orbfe_last_error());
}

// cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) + `d0 < d1*0.7`  (Frame::ComputeStereoFishEyeMatches,
// Frame.cc:1553-1562): match[i] = accepted train row or -1.
inline void KnnRatioMatch(const cv::Mat& query, const cv::Mat& train, std::vector<int32_t>& match,
                          std::vector<int32_t>& idx2, std::vector<int32_t>& dist2) {
    match.assign(query.rows, -1);
    idx2.assign((size_t)query.rows * 2, -1);
    dist2.assign((size_t)query.rows * 2, -1);
    if (orbfe_knn2(query.ptr(), query.rows, train.ptr(), train.rows, 0, idx2.data(), dist2.data(), match.data(),
                   device()) != ORBFE_OK)
        throw std::runtime_error(std::string("KnnRatioMatch (B200): ") + orbfe_last_error());
}

// The geometry step that follows KnnRatioMatch in Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1560-1587):
// KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:439-515) for ALL ratio-test survivors in one
// call.  params1 / params2 = the two cameras' mvParameters (8 floats), precision = GetPrecision(); R12 row major
// (Frame::mRlr.data() is column major in Eigen: pass mRlr.transpose().data() or copy), t12 = mtlr; pairs[i] =
// (index into keysL, index into keysR); levelSigma2 = mvLevelSigma2.  depth[i] / p3D[3i..] as the reference returns
// them per match; the caller keeps :1580-1586 (depth > 0.0001f -> mvLeftToRightMatch, mvStereo3Dpoints, mvDepth).
inline void TriangulateFisheyeMatches(const float params1[8], float precision1, const float params2[8], float precision2,
                                      const float R12[9], const float t12[3], const std::vector<cv::KeyPoint>& keysL,
                                      const std::vector<cv::KeyPoint>& keysR,
                                      const std::vector<std::pair<int, int> >& pairs, const std::vector<float>& levelSigma2,
                                      std::vector<float>& depth, std::vector<float>& p3D) {
    const size_t n = pairs.size();
    std::vector<float> pt1(2 * n), pt2(2 * n), s1(n), s2(n);
    for (size_t i = 0; i < n; i++) {
        const cv::KeyPoint &a = keysL[pairs[i].first], &b = keysR[pairs[i].second];
        pt1[2 * i] = a.pt.x; pt1[2 * i + 1] = a.pt.y;
        pt2[2 * i] = b.pt.x; pt2[2 * i + 1] = b.pt.y;
        s1[i] = levelSigma2[a.octave];
        s2[i] = levelSigma2[b.octave];
    }
    depth.assign(n, -1.f);
    p3D.assign(3 * n, 0.f);
    if (orbfe_kb8_triangulate_matches(params1, precision1, params2, precision2, R12, t12, pt1.data(), pt2.data(), s1.data(),
                                      s2.data(), (int)n, depth.data(), p3D.data(), device()) != ORBFE_OK)
        throw std::runtime_error(std::string("TriangulateFisheyeMatches (B200): ") + orbfe_last_error());
}

}  // namespace b200
}  // namespace ORB_SLAM3

#endif
