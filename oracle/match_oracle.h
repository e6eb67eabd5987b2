// oracle/match_oracle.h -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
//
// CPU restatement of the reference's Hamming-matching path.  The reference functions work on
// Frame / MapPoint / KeyFrame object graphs (Eigen, Sophus, DBoW2 - none of which exist in
// this image), so their files cannot be compiled as a whole; each function below follows the
// control flow of the cited lines on plain arrays marshalled at the boundary that
// SURVEY.md section 8(b) defines (projected coordinates and per-point flags are inputs).
//   ORBmatcher::DescriptorDistance            /root/reference/src/ORBmatcher.cc:2384-2404
//   Frame::AssignFeaturesToGrid / PosInGrid   src/Frame.cc:469-504, 962-978
//   Frame::GetFeaturesInArea                  src/Frame.cc:859-951
//   ORBmatcher::SearchByProjection(F, MPs)    src/ORBmatcher.cc:46-250   (Nleft == -1 branch)
//   ORBmatcher::SearchByProjection(Cur, Last) src/ORBmatcher.cc:1951-2185 (Nleft == -1 branch)
//   ORBmatcher::SearchByProjection(F, KF, ..) src/ORBmatcher.cc:2197-2325
//   ORBmatcher::ComputeThreeMaxima            src/ORBmatcher.cc:2336-2378
//   Frame::ComputeStereoMatches               src/Frame.cc:1102-1358
//   Frame::ComputeStereoFishEyeMatches        src/Frame.cc:1530-1587 (kNN-2 + 0.7 ratio part)
// Parity status: the reference has no tests for this path; PINNED against the reference's own function
// bodies compiled verbatim into oracle/_ref/libref_orbmatcher.so (ref_build.sh, ref_slices.py, refshim/)
// by tests/test_oracle_match_vs_ref.py -- all of the above except the fisheye kNN part, which is pinned
// against cv2 4.13 BFMatcher (tests/test_oracle_cvprims.py).
#pragma once
#include <cstdint>
#include <vector>

#include "orb_oracle.h"

namespace match_oracle {

using orb_oracle::OrbKp;

enum { GRID_COLS = 64, GRID_ROWS = 48 };  // reference include/Frame.h:44-45
enum { TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30 };  // src/ORBmatcher.cc:36-38

// The part of ORB_SLAM3::Frame that the matchers read (mono / rectified stereo / RGB-D
// layout, i.e. Nleft == -1).
struct FrameView {
    int N = 0;
    const OrbKp* keys = nullptr;    // mvKeysUn
    const float* uright = nullptr;  // mvuRight (nullptr => monocular, all -1)
    const uint8_t* desc = nullptr;  // mDescriptors, N x 32
    float minX = 0, minY = 0, maxX = 0, maxY = 0;  // mnMinX .. mnMaxY
    float gridWInv = 0, gridHInv = 0;              // mfGridElementWidthInv / HeightInv
    const float* scaleFactors = nullptr;           // mvScaleFactors
    int nlevels = 0;
    std::vector<int> grid[GRID_COLS][GRID_ROWS];   // mGrid
    void assign_features_to_grid();
    std::vector<int> features_in_area(float x, float y, float r, int minLevel, int maxLevel) const;
};

int descriptor_distance(const uint8_t* a, const uint8_t* b);

// One projected map point as the matcher sees it.
struct ProjPoint {
    float u, v;        // projection into the frame (mTrackProjX/Y or project(x3Dc))
    float ur;          // predicted right-image u (mTrackProjXR, or u - mbf*invz)
    float radius;      // search window half-size, r*scaleFactor[level] already applied
    int minLevel, maxLevel;  // GetFeaturesInArea level gate
    float angle;       // keypoint angle on the source side (rotation histogram)
    uint8_t valid;     // passes every early `continue` of the reference loop
    uint8_t blocks;    // its MapPoint::Observations() > 0 (a claim blocks later points)
};

struct SearchParams {
    int mode;          // 0: (F, MapPoints) best+second+ratio; 1: (Cur, Last); 2: (F, KF)
    int thAccept;      // TH_HIGH for modes 0/1, ORBdist for mode 2
    float nnratio;     // mfNNratio (mode 0)
    int checkOrientation;  // mbCheckOrientation (modes 1/2)
};

// claimed[i] != 0 <=> F.mvpMapPoints[i] && Observations()>0 on entry (mode 2: non-null).
// assigned[i] (in/out) = index of the ProjPoint now held by keypoint i, or -1 / untouched.
// best_idx[j] = keypoint accepted for point j (before rotation culling) or -1.
int search_by_projection(FrameView& F, const std::vector<ProjPoint>& pts, const uint8_t* pdesc,
                         const SearchParams& prm, const uint8_t* claimed, int* assigned,
                         int* best_idx, int* best_dist);

// The same two overloads for a fisheye stereo frame (Nleft != -1): src/ORBmatcher.cc:46-240 incl. the
// right-camera branch :171-237 and the mvLeftToRightMatch / mvRightToLeftMatch partner writes
// (:159-163, :215-219), and :1951-2185 incl. the right-camera branch :2090-2155.  FL holds mvKeys /
// left grid / descriptor rows [0,Nleft), FR holds mvKeysRight / right grid / rows [Nleft,N).
// claimed / assigned are indexed like F.mvpMapPoints: [0,Nleft) left, [Nleft,N) right.
int search_by_projection_fisheye(FrameView& FL, FrameView& FR, const int* l2r, const int* r2l,
                                 const std::vector<ProjPoint>& ptsL, const std::vector<ProjPoint>& ptsR,
                                 const uint8_t* pdesc, const SearchParams& prm, const uint8_t* claimed,
                                 int* assigned, int* best_idx_l, int* best_idx_r);

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
// src/ORBmatcher.cc:735-891.  F1 needs keys / desc only; F2 is searched through its grid.
// prevMatched: n1 x 2 floats (in/out), matches12: n1 ints (out).  Returns nmatches.
int search_for_initialization(const FrameView& F1, FrameView& F2, float* prevMatched, int windowSize,
                              float nnratio, bool checkOrientation, int* matches12);

// ---- keyframe-side searches --------------------------------------------------
// The inner loop shared by ORBmatcher::Fuse (src/ORBmatcher.cc:1326-1534 and :1536-1688),
// SearchBySim3 (:1690-1940) and the Sim3 SearchByProjection overloads (:496-733): KeyFrame::
// GetFeaturesInArea (src/KeyFrame.cc:843-892, same cells and order as Frame's) followed by the
// `kpLevel < nPredictedLevel-1 || kpLevel > nPredictedLevel` filter, optionally Fuse's reprojection
// gate (:1436-1461: chi2 7.8 with a stereo coordinate, 5.99 without, on e2 * mvInvLevelSigma2[level]),
// and a strict-`<` Hamming argmin.  No state is carried from one point to the next.
// best_idx[j] = keypoint index or -1 (none within thAccept), best_dist[j] = best distance (256 = none).
struct WindowParams {
    int thAccept;                   // TH_LOW (Fuse), TH_HIGH (SearchBySim3)
    int fuseGate;                   // 1 = apply Fuse's chi2 gate
    const float* invLevelSigma2;    // KeyFrame::mvInvLevelSigma2 (gate only)
    int nlevels;
};
void search_window(const FrameView& KF, const std::vector<ProjPoint>& pts, const uint8_t* pdesc,
                   const WindowParams& prm, int* best_idx, int* best_dist);

// ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1690-1940) after the caller's projections: pts12[i1] =
// map point of KF1 slot i1 projected into KF2 (valid = has a good, not yet matched point that passes
// the depth / image / distance checks), pts21 likewise; desc1 / desc2 = GetDescriptor() of those points.
// match12[i1] = KF2 keypoint index when both directions agree (:1925-1937), else -1.  Returns nFound.
int search_by_sim3(const FrameView& KF1, const FrameView& KF2, const std::vector<ProjPoint>& pts12,
                   const uint8_t* desc1, const std::vector<ProjPoint>& pts21, const uint8_t* desc2,
                   int thAccept, int* match12);

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:438-529) for one map point: desc = its N observed
// descriptors in vDescriptors order.  Returns BestIdx (:507-521), -1 when N == 0.
int distinctive_descriptor(const uint8_t* desc, int N);

// Frame::ComputeStereoMatches on two extractor pyramids.
struct PyrLevelView {
    const uint8_t* roi;
    int w, h, step;
};
void compute_stereo_matches(const OrbKp* keysL, const uint8_t* descL, int N, const OrbKp* keysR,
                            const uint8_t* descR, int Nr, const PyrLevelView* pyrL,
                            const PyrLevelView* pyrR, const float* scaleFactors,
                            const float* invScaleFactors, float mbf, float mb, float* uRight,
                            float* depth);

// kNN-2 + `d0 < d1*0.7` of ComputeStereoFishEyeMatches: match[i] = train index or -1.
void fisheye_ratio_matches(const uint8_t* q, int nq, const uint8_t* t, int nt, int* match,
                           int* idx2, int* dist2);

}  // namespace match_oracle
