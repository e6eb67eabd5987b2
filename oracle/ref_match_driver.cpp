// oracle/ref_match_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the reference's OWN matcher functions: ORBmatcher::SearchByProjection (three
// Frame overloads), ORBmatcher::SearchForInitialization, ORBmatcher::DescriptorDistance,
// Frame::GetFeaturesInArea / AssignFeaturesToGrid / PosInGrid, Frame::ComputeStereoMatches and
// MapPoint::PredictScale.  Their bodies are cut out of /root/reference/src at build time by
// oracle/ref_slices.py and compiled verbatim against oracle/refshim + oracle/cvshim (ref_build.sh);
// this file only builds the Frame / MapPoint objects they walk from plain arrays and flattens the result.
// Used by tests/test_oracle_match_vs_ref.py to pin oracle/match_oracle.cpp.
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "refshim.h"
#include "ORBmatcher.h"
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

using namespace ORB_SLAM3;

std::vector<RefAction> ORB_SLAM3::g_refActions;
float Frame::mfGridElementWidthInv = 0, Frame::mfGridElementHeightInv = 0;
float Frame::mnMinX = 0, Frame::mnMaxX = 0, Frame::mnMinY = 0, Frame::mnMaxY = 0;

namespace {

struct RefFrame {
    Frame F;
    Pinhole cam;
    std::vector<std::unique_ptr<MapPoint>> own;   // map points held by F.mvpMapPoints (slot order, may be null)
    std::vector<MapPoint*> initial;               // F.mvpMapPoints before a search
    KeyFrame kf;                                  // keyframe view of the same data (relocalisation overload)
    std::unique_ptr<ORBextractor> exL, exR;
};

class Matcher : public ORBmatcher {
   public:
    using ORBmatcher::ORBmatcher;
};

cv::Mat wrap_desc(const uint8_t* d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8UC1);
    if (n > 0) memcpy(m.data, d, (size_t)n * 32);
    return m;
}

// slot_out[i]: index (into `cands`) of the map point now held by slot i; -2 = the slot still holds what it
// held before the search; -1 = the slot is now NULL but was not before.
void flatten(RefFrame* rf, const std::vector<MapPoint*>& cands, int* slot_out) {
    std::map<MapPoint*, int> index;
    for (size_t i = 0; i < cands.size(); i++)
        if (cands[i] && !index.count(cands[i])) index[cands[i]] = (int)i;
    for (int i = 0; i < rf->F.N; i++) {
        MapPoint* p = rf->F.mvpMapPoints[i];
        if (p == rf->initial[i]) slot_out[i] = -2;
        else if (!p) slot_out[i] = -1;
        else slot_out[i] = index.count(p) ? index[p] : -3;
    }
}

}  // namespace

extern "C" {

void refm_set_bounds(float minX, float minY, float maxX, float maxY, float gwInv, float ghInv) {
    Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = gwInv; Frame::mfGridElementHeightInv = ghInv;
}

// keys/desc: the left (or only) view; keysR/descR with nR >= 0: the right view of a fisheye stereo frame
// (Frame::Nleft != -1).  uRight may be null.  The grid is assigned with the reference's own code.
void* refm_frame_create(const cv::KeyPoint* keys, int n, const uint8_t* desc, const float* uRight,
                        const cv::KeyPoint* keysR, int nR, const uint8_t* descR, const int* l2r, const int* r2l,
                        const float* scaleFactors, int nlevels, float logScaleFactor, float mb, float mbf) {
    RefFrame* rf = new RefFrame();
    Frame& F = rf->F;
    F.mpCamera = &rf->cam;
    F.mvKeys.assign(keys, keys + n);
    F.mvKeysUn = F.mvKeys;
    F.mb = mb; F.mbf = mbf;
    F.mnScaleLevels = nlevels;
    F.mfLogScaleFactor = logScaleFactor;
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.mvInvScaleFactors.resize(nlevels);
    for (int i = 0; i < nlevels; i++) F.mvInvScaleFactors[i] = 1.0f / scaleFactors[i];
    if (nR >= 0) {
        F.Nleft = n; F.Nright = nR; F.N = n + nR;
        F.mvKeysRight.assign(keysR, keysR + nR);
        std::vector<uint8_t> all((size_t)(n + nR) * 32 + 32);
        if (n) memcpy(all.data(), desc, (size_t)n * 32);
        if (nR) memcpy(all.data() + (size_t)n * 32, descR, (size_t)nR * 32);
        F.mDescriptors = wrap_desc(all.data(), n + nR);
        F.mvLeftToRightMatch.assign(l2r, l2r + n);
        F.mvRightToLeftMatch.assign(r2l, r2l + nR);
        F.mvuRight.assign(F.N, -1.f);
    } else {
        F.N = n;
        F.mDescriptors = wrap_desc(desc, n);
        if (uRight) F.mvuRight.assign(uRight, uRight + n); else F.mvuRight.assign(n, -1.f);
    }
    F.mvDepth.assign(F.N, -1.f);
    F.mvpMapPoints.assign(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    rf->own.resize(F.N);
    F.AssignFeaturesToGrid();
    rf->initial = F.mvpMapPoints;
    return rf;
}
void refm_frame_destroy(void* h) { delete (RefFrame*)h; }
void refm_frame_pose(void* h, float tx, float ty, float tz) { ((RefFrame*)h)->F.mTcw = Sophus::SE3f(Eigen::Vector3f(tx, ty, tz)); }
void refm_frame_trl(void* h, float tx, float ty, float tz) { ((RefFrame*)h)->F.mTrl = Sophus::SE3f(Eigen::Vector3f(tx, ty, tz)); }

// Per slot: has[i] != 0 -> the slot holds a map point with nobs[i] observations, world position xyz[3i..],
// descriptor desc[32i..], distance range [minDist, maxDist] (MapPoint::mfMinDistance / mfMaxDistance).
void refm_frame_mappoints(void* h, const uint8_t* has, const int* nobs, const float* xyz, const uint8_t* desc,
                          const uint8_t* outlier, const uint8_t* bad, const float* minDist, const float* maxDist) {
    RefFrame* rf = (RefFrame*)h;
    Frame& F = rf->F;
    for (int i = 0; i < F.N; i++) {
        rf->own[i].reset();
        F.mvpMapPoints[i] = nullptr;
        if (outlier) F.mvbOutlier[i] = outlier[i] != 0;
        if (!has[i]) continue;
        MapPoint* p = new MapPoint();
        rf->own[i].reset(p);
        p->nObs = nobs ? nobs[i] : 1;
        if (xyz) p->pos = Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        p->desc = wrap_desc(desc ? desc + 32 * (size_t)i : nullptr, desc ? 1 : 0);
        p->bad = bad ? bad[i] != 0 : false;
        if (minDist) p->mfMinDistance = minDist[i];
        if (maxDist) p->mfMaxDistance = maxDist[i];
        F.mvpMapPoints[i] = p;
    }
    rf->initial = F.mvpMapPoints;
}

int refm_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma = wrap_desc(a, 1), mb = wrap_desc(b, 1);
    return ORBmatcher::DescriptorDistance(ma, mb);
}

int refm_features_in_area(void* h, float x, float y, float r, int minLevel, int maxLevel, int bRight, int* out, int cap) {
    std::vector<size_t> v = ((RefFrame*)h)->F.GetFeaturesInArea(x, y, r, minLevel, maxLevel, bRight != 0);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}

int refm_predict_scale(void* h, float maxDistance, float dist) {
    MapPoint p;
    p.mfMaxDistance = maxDistance;
    return p.PredictScale(dist, &((RefFrame*)h)->F);
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
int refm_search_mappoints(void* h, int m, const uint8_t* inView, const uint8_t* inViewR, const float* depth,
                          const uint8_t* bad, const int* nobs, const float* projX, const float* projY,
                          const float* projXR, const float* projYR, const int* level, const int* levelR,
                          const float* viewCos, const float* viewCosR, const uint8_t* desc, float th, int bFar,
                          float thFar, float nnratio, int* slot_out) {
    RefFrame* rf = (RefFrame*)h;
    std::vector<std::unique_ptr<MapPoint>> pts(m);
    std::vector<MapPoint*> vp(m);
    for (int i = 0; i < m; i++) {
        MapPoint* p = new MapPoint();
        pts[i].reset(p); vp[i] = p;
        p->mbTrackInView = inView[i] != 0;
        p->mbTrackInViewR = inViewR ? inViewR[i] != 0 : false;
        p->mTrackDepth = depth ? depth[i] : 1.f;
        p->bad = bad ? bad[i] != 0 : false;
        p->nObs = nobs ? nobs[i] : 1;
        p->mTrackProjX = projX[i]; p->mTrackProjY = projY[i];
        p->mTrackProjXR = projXR ? projXR[i] : 0.f; p->mTrackProjYR = projYR ? projYR[i] : 0.f;
        p->mnTrackScaleLevel = level[i]; p->mnTrackScaleLevelR = levelR ? levelR[i] : -1;
        p->mTrackViewCos = viewCos[i]; p->mTrackViewCosR = viewCosR ? viewCosR[i] : 1.f;
        p->desc = wrap_desc(desc + 32 * (size_t)i, 1);
    }
    Matcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(rf->F, vp, th, bFar != 0, thFar);
    flatten(rf, vp, slot_out);
    return n;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)
int refm_search_lastframe(void* cur, void* last, float th, int bMono, float nnratio, int checkOri, int* slot_out) {
    RefFrame *rc = (RefFrame*)cur, *rl = (RefFrame*)last;
    Matcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByProjection(rc->F, rl->F, th, bMono != 0);
    flatten(rc, rl->F.mvpMapPoints, slot_out);
    return n;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist)
int refm_search_keyframe(void* cur, void* kfFrame, const uint8_t* alreadyFound, float th, int ORBdist,
                         float nnratio, int checkOri, int* slot_out) {
    RefFrame *rc = (RefFrame*)cur, *rk = (RefFrame*)kfFrame;
    rk->kf.mvpMapPoints = rk->F.mvpMapPoints;
    rk->kf.mvKeysUn = rk->F.mvKeysUn;
    std::set<MapPoint*> found;
    for (int i = 0; i < rk->F.N; i++)
        if (alreadyFound && alreadyFound[i] && rk->F.mvpMapPoints[i]) found.insert(rk->F.mvpMapPoints[i]);
    Matcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByProjection(rc->F, &rk->kf, found, th, ORBdist);
    flatten(rc, rk->F.mvpMapPoints, slot_out);
    return n;
}

// ---- keyframe-side searches ---------------------------------------------------------------------------
// The RefFrame's data seen as a KeyFrame, filled the way KeyFrame::KeyFrame(Frame&, ...) does
// (src/KeyFrame.cc:44-76): keys, descriptors, grid copied cell by cell, bounds truncated to int.
static KeyFrame* as_keyframe(RefFrame* rf) {
    KeyFrame& K = rf->kf;
    const Frame& F = rf->F;
    K.N = F.N; K.NLeft = F.Nleft;
    K.mvKeys = F.mvKeys; K.mvKeysUn = F.mvKeysUn; K.mvKeysRight = F.mvKeysRight;
    K.mvuRight = F.mvuRight; K.mDescriptors = F.mDescriptors;
    K.mnScaleLevels = F.mnScaleLevels; K.mfLogScaleFactor = F.mfLogScaleFactor;
    K.mvScaleFactors = F.mvScaleFactors;
    K.mvInvLevelSigma2.resize(F.mnScaleLevels);
    K.mvLevelSigma2.resize(F.mnScaleLevels);
    for (int i = 0; i < F.mnScaleLevels; i++) {   // ORBextractor.cc:426-437
        K.mvLevelSigma2[i] = F.mvScaleFactors[i] * F.mvScaleFactors[i];
        K.mvInvLevelSigma2[i] = 1.0f / K.mvLevelSigma2[i];
    }
    K.mnMinX = Frame::mnMinX; K.mnMinY = Frame::mnMinY; K.mnMaxX = Frame::mnMaxX; K.mnMaxY = Frame::mnMaxY;
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    K.mGrid.assign(K.mnGridCols, std::vector<std::vector<size_t>>(K.mnGridRows));
    for (int i = 0; i < K.mnGridCols; i++)
        for (int j = 0; j < K.mnGridRows; j++) K.mGrid[i][j] = F.mGrid[i][j];
    K.fx = rf->cam.fx; K.fy = rf->cam.fy; K.cx = rf->cam.cx; K.cy = rf->cam.cy; K.mbf = F.mbf;
    K.mpCamera = &rf->cam; K.mpCamera2 = &rf->cam;   // Fuse(bRight) reads mpCamera2; the BoW / triangulation drivers reset it
    K.mTcw = F.mTcw;
    K.mvpMapPoints = F.mvpMapPoints;
    for (int i = 0; i < F.N; i++)
        if (K.mvpMapPoints[i]) K.mvpMapPoints[i]->id = 1000000 + i;
    return &K;
}

static std::vector<std::unique_ptr<MapPoint>> make_points(int m, const uint8_t* has, const uint8_t* bad, const uint8_t* inKF,
                                                          const float* xyz, const float* normal, const float* minDist,
                                                          const float* maxDist, const int* nobs, const uint8_t* desc,
                                                          std::vector<MapPoint*>& vp) {
    std::vector<std::unique_ptr<MapPoint>> pts(m);
    vp.assign(m, nullptr);
    for (int i = 0; i < m; i++) {
        if (has && !has[i]) continue;
        MapPoint* p = new MapPoint();
        pts[i].reset(p); vp[i] = p;
        p->id = i;
        p->bad = bad ? bad[i] != 0 : false;
        p->inKF = inKF ? inKF[i] != 0 : false;
        p->pos = Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        if (normal) p->normal = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p->mfMinDistance = minDist[i]; p->mfMaxDistance = maxDist[i];
        p->nObs = nobs ? nobs[i] : 1;
        p->desc = wrap_desc(desc + 32 * (size_t)i, 1);
    }
    return pts;
}

static int dump_actions(int* actions, int cap) {
    int n = 0;
    for (const RefAction& a : g_refActions) {
        if (n < cap) { actions[3 * n] = a.kind; actions[3 * n + 1] = a.a; actions[3 * n + 2] = a.b; }
        n++;
    }
    return n;
}

void refm_frame_camera(void* h, float fx, float fy, float cx, float cy) {
    RefFrame* rf = (RefFrame*)h;
    rf->cam.fx = fx; rf->cam.fy = fy; rf->cam.cx = cx; rf->cam.cy = cy;
}

int refm_kf_predict_scale(void* h, float maxDistance, float dist) {
    MapPoint p;
    p.mfMaxDistance = maxDistance;
    return p.PredictScale(dist, as_keyframe((RefFrame*)h));
}

// ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight).  Graph updates come back as a list of
// (kind, a, b): 1 = point a Replace(point b), 2 = point a AddObservation(idx b), 3 = AddMapPoint(point a, idx b);
// candidate points have id = their index, keyframe slot points id = 1000000 + slot.
int refm_fuse(void* h, int m, const uint8_t* has, const uint8_t* bad, const uint8_t* inKF, const float* xyz,
              const float* normal, const float* minDist, const float* maxDist, const int* nobs, const uint8_t* desc,
              float th, int* actions, int cap, int* nActions) {
    RefFrame* rf = (RefFrame*)h;
    KeyFrame* kf = as_keyframe(rf);
    std::vector<MapPoint*> vp;
    auto own = make_points(m, has, bad, inKF, xyz, normal, minDist, maxDist, nobs, desc, vp);
    g_refActions.clear();
    Matcher matcher(0.6f, true);
    const int n = matcher.Fuse(kf, vp, th, false);
    *nActions = dump_actions(actions, cap);
    return n;
}

// ORBmatcher::Fuse(KeyFrame*, Sim3f& Scw, vpPoints, th, vpReplacePoint): replace_out[i] = keyframe slot whose
// point replaces candidate i, or -1.
int refm_fuse_sim3(void* h, float s, float tx, float ty, float tz, int m, const uint8_t* bad, const float* xyz,
                   const float* normal, const float* minDist, const float* maxDist, const uint8_t* desc, float th,
                   int* replace_out, int* actions, int cap, int* nActions) {
    RefFrame* rf = (RefFrame*)h;
    KeyFrame* kf = as_keyframe(rf);
    std::vector<MapPoint*> vp;
    auto own = make_points(m, nullptr, bad, nullptr, xyz, normal, minDist, maxDist, nullptr, desc, vp);
    std::vector<MapPoint*> repl(m, nullptr);
    Sophus::Sim3f Scw(s, Eigen::Vector3f(tx, ty, tz));
    g_refActions.clear();
    Matcher matcher(0.6f, true);
    const int n = matcher.Fuse(kf, Scw, vp, th, repl);
    for (int i = 0; i < m; i++) replace_out[i] = repl[i] ? repl[i]->id - 1000000 : -1;
    *nActions = dump_actions(actions, cap);
    return n;
}

// ORBmatcher::SearchByProjection(KeyFrame*, Sim3f& Scw, vpPoints, vpMatched, th, ratioHamming):
// matched_in[i] != 0 <=> vpMatched[i] non-NULL on entry; slot_out[i] = index of the candidate now in vpMatched[i],
// -2 = unchanged.
int refm_search_kf_sim3(void* h, float s, float tx, float ty, float tz, int m, const uint8_t* bad, const float* xyz,
                        const float* normal, const float* minDist, const float* maxDist, const uint8_t* desc,
                        const uint8_t* matched_in, int th, float ratioHamming, int* slot_out) {
    RefFrame* rf = (RefFrame*)h;
    KeyFrame* kf = as_keyframe(rf);
    std::vector<MapPoint*> vp;
    auto own = make_points(m, nullptr, bad, nullptr, xyz, normal, minDist, maxDist, nullptr, desc, vp);
    std::vector<std::unique_ptr<MapPoint>> placeholders;
    std::vector<MapPoint*> matched(kf->N, nullptr);
    for (int i = 0; i < kf->N; i++)
        if (matched_in[i]) { placeholders.emplace_back(new MapPoint()); placeholders.back()->id = -7; matched[i] = placeholders.back().get(); }
    const std::vector<MapPoint*> before = matched;
    Sophus::Sim3f Scw(s, Eigen::Vector3f(tx, ty, tz));
    Matcher matcher(0.6f, true);
    const int n = matcher.SearchByProjection(kf, Scw, vp, matched, th, ratioHamming);
    for (int i = 0; i < kf->N; i++) slot_out[i] = matched[i] == before[i] ? -2 : matched[i]->id;
    return n;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th): pre12[i1] = KF2 slot whose point is already in
// vpMatches12[i1] (-1 = NULL); out12[i1] = KF2 slot of the point in vpMatches12[i1] afterwards, or -1.
int refm_search_by_sim3(void* h1, void* h2, float s, float tx, float ty, float tz, const int* pre12, float th, int* out12) {
    RefFrame *r1 = (RefFrame*)h1, *r2 = (RefFrame*)h2;
    KeyFrame* k1 = as_keyframe(r1);
    KeyFrame* k2 = as_keyframe(r2);
    for (int i = 0; i < k2->N; i++)
        if (k2->mvpMapPoints[i]) { k2->mvpMapPoints[i]->id = 2000000 + i; k2->mvpMapPoints[i]->idxInOtherKF = i; }
    std::vector<MapPoint*> m12(k1->N, nullptr);
    for (int i = 0; i < k1->N; i++)
        if (pre12[i] >= 0) m12[i] = k2->mvpMapPoints[pre12[i]];
    Sophus::Sim3f S12(s, Eigen::Vector3f(tx, ty, tz));
    Matcher matcher(0.75f, true);
    const int n = matcher.SearchBySim3(k1, k2, m12, S12, th);
    for (int i = 0; i < k1->N; i++) out12[i] = m12[i] ? m12[i]->id - 2000000 : -1;
    return n;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
int refm_search_init(void* f1, void* f2, float* prevMatched, int* matches12, int windowSize, float nnratio, int checkOri) {
    RefFrame *r1 = (RefFrame*)f1, *r2 = (RefFrame*)f2;
    const int n1 = (int)r1->F.mvKeysUn.size();
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prevMatched[2 * i], prevMatched[2 * i + 1]);
    std::vector<int> m12;
    Matcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForInitialization(r1->F, r2->F, prev, m12, windowSize);
    for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prevMatched[2 * i] = prev[i].x; prevMatched[2 * i + 1] = prev[i].y; }
    return n;
}

// Frame::ComputeStereoMatches() on a rectified pair: both images go through the reference's own
// ORBextractor (src/ORBextractor.cc, compiled into this library too).  Returns N (left keypoints);
// keys/desc of both views are written out so that the restatement can be fed the same inputs.
int refm_stereo(const uint8_t* imgL, const uint8_t* imgR, int rows, int cols, int nfeatures, float scaleFactor,
                int nlevels, int iniTh, int minTh, float mbf, float mb, cv::KeyPoint* keysL, uint8_t* descL,
                cv::KeyPoint* keysR, uint8_t* descR, int cap, int* nR_out, float* uRight, float* depth) {
    RefFrame rf;
    Frame& F = rf.F;
    rf.exL.reset(new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh));
    rf.exR.reset(new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh));
    F.mpORBextractorLeft = rf.exL.get(); F.mpORBextractorRight = rf.exR.get();
    cv::Mat L(rows, cols, CV_8UC1, (void*)imgL, (size_t)cols), R(rows, cols, CV_8UC1, (void*)imgR, (size_t)cols);
    std::vector<int> lap = {0, 0};
    (*rf.exL)(L, cv::Mat(), F.mvKeys, F.mDescriptors, lap);
    (*rf.exR)(R, cv::Mat(), F.mvKeysRight, F.mDescriptorsRight, lap);
    F.N = (int)F.mvKeys.size();
    F.mvScaleFactors = rf.exL->GetScaleFactors();
    F.mvInvScaleFactors = rf.exL->GetInverseScaleFactors();
    F.mbf = mbf; F.mb = mb;
    const int nR = (int)F.mvKeysRight.size();
    *nR_out = nR;
    if (F.N > cap || nR > cap) return -1;
    F.ComputeStereoMatches();
    for (int i = 0; i < F.N; i++) {
        keysL[i] = F.mvKeys[i]; memcpy(descL + 32 * (size_t)i, F.mDescriptors.ptr(i), 32);
        uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i];
    }
    for (int i = 0; i < nR; i++) { keysR[i] = F.mvKeysRight[i]; memcpy(descR + 32 * (size_t)i, F.mDescriptorsRight.ptr(i), 32); }
    return F.N;
}


// MapPoint::ComputeDistinctiveDescriptors for a batch of map points.  Point p is observed by nkf[p] keyframes (allocated
// as one array, so that the std::map<KeyFrame*, ...> of observations iterates in index order); keyframe k of the point
// contributes rows[k] (1 or 2: left only, or left + right index) descriptors, bad keyframes are skipped (:453).
// desc = all rows in that order; out = the chosen 32-byte mDescriptor per point (untouched when nothing was chosen).
void refm_distinctive(const uint8_t* desc, const int* kfStart, const int* rows, const uint8_t* kfBad, int nPoints,
                      uint8_t* out) {
    size_t row = 0;
    for (int p = 0; p < nPoints; p++) {
        const int nk = kfStart[p + 1] - kfStart[p];
        std::vector<KeyFrame> kfs(nk);
        MapPoint mp;
        for (int k = 0; k < nk; k++) {
            const int g = kfStart[p] + k, r = rows[g];
            kfs[k].mDescriptors = wrap_desc(desc + 32 * row, r);
            kfs[k].mbBadKF = kfBad[g] != 0;
            mp.mObservations[&kfs[k]] = std::make_tuple(0, r == 2 ? 1 : -1);
            row += r;
        }
        mp.ComputeDistinctiveDescriptors();
        if (!mp.mDescriptor.empty()) memcpy(out + 32 * (size_t)p, mp.mDescriptor.ptr(), 32);
    }
}

// ---- bag of words: DBoW2 itself (include/ORBVocabulary.h:30-31 typedef) ------------------------------------
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;

void* refd_voc_load(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }   // System.cc:105
    return v;
}
void refd_voc_destroy(void* h) { delete (ORBVocabulary*)h; }
int refd_voc_size(void* h) { return (int)((ORBVocabulary*)h)->size(); }

static std::vector<cv::Mat> to_descriptor_vector(const uint8_t* desc, int n, cv::Mat& hold) {
    hold = wrap_desc(desc, n);
    std::vector<cv::Mat> v;                       // Converter::toDescriptorVector, src/Converter.cc:27-39
    for (int j = 0; j < n; j++) v.push_back(hold.row(j));
    return v;
}

// per feature: word id (public transform), its weight, and the node `levelsup` levels above the word
void refd_voc_transform_features(void* h, const uint8_t* desc, int n, int levelsup, int* word, double* weight, int* nid) {
    ORBVocabulary* voc = (ORBVocabulary*)h;
    cv::Mat hold;
    std::vector<cv::Mat> v = to_descriptor_vector(desc, n, hold);
    for (int i = 0; i < n; i++) {
        const DBoW2::WordId w = voc->transform(v[i]);
        word[i] = (int)w; weight[i] = voc->getWordWeight(w); nid[i] = (int)voc->getParentNode(w, levelsup);
    }
}

static void flatten_fv(const DBoW2::FeatureVector& fv, int* nNodes, unsigned* nodes, int* start, unsigned* feat) {
    int j = 0, p = 0;
    for (auto& e : fv) {
        nodes[j] = e.first; start[j] = p;
        for (unsigned f : e.second) feat[p++] = f;
        j++;
    }
    start[j] = p;
    *nNodes = j;
}

// Frame::ComputeBoW (src/Frame.cc:984-998): transform(vCurrentDesc, mBowVec, mFeatVec, 4)
void refd_voc_transform(void* h, const uint8_t* desc, int n, int levelsup, int* nWords, unsigned* ids, double* values,
                        int* nNodes, unsigned* nodes, int* start, unsigned* feat) {
    ORBVocabulary* voc = (ORBVocabulary*)h;
    cv::Mat hold;
    std::vector<cv::Mat> v = to_descriptor_vector(desc, n, hold);
    DBoW2::BowVector bow;
    DBoW2::FeatureVector fv;
    voc->transform(v, bow, fv, levelsup);
    int i = 0;
    for (auto& e : bow) { ids[i] = e.first; values[i] = e.second; i++; }
    *nWords = i;
    flatten_fv(fv, nNodes, nodes, start, feat);
}

// ComputeBoW on a RefFrame (frame and keyframe view) with the given vocabulary
void refm_frame_compute_bow(void* fh, void* vh, int levelsup) {
    RefFrame* rf = (RefFrame*)fh;
    ORBVocabulary* voc = (ORBVocabulary*)vh;
    std::vector<cv::Mat> v;
    for (int j = 0; j < rf->F.N; j++) v.push_back(rf->F.mDescriptors.row(j));
    voc->transform(v, rf->F.mBowVec, rf->F.mFeatVec, levelsup);
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches):
// out[i] (size F.N) = keyframe slot whose map point is now in vpMapPointMatches[i], or -1.
int refm_search_by_bow_kf_f(void* kfh, void* fh, float nnratio, int checkOri, int* out) {
    RefFrame *rk = (RefFrame*)kfh, *rf = (RefFrame*)fh;
    KeyFrame* kf = as_keyframe(rk);
    kf->mFeatVec = rk->F.mFeatVec;
    kf->mpCamera2 = (rk->F.Nleft != -1) ? &rk->cam : nullptr;     // :357, :380: stereo-fisheye keyframes have a second camera
    rf->F.mpCamera2 = (rf->F.Nleft != -1) ? &rf->cam : nullptr;
    std::vector<MapPoint*> matches;
    Matcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(kf, rf->F, matches);
    for (int i = 0; i < rf->F.N; i++) out[i] = matches[i] ? matches[i]->id - 1000000 : -1;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12):
// out[i1] = KF2 slot of the point in vpMatches12[i1], or -1.
int refm_search_by_bow_kf_kf(void* h1, void* h2, float nnratio, int checkOri, int* out) {
    RefFrame *r1 = (RefFrame*)h1, *r2 = (RefFrame*)h2;
    KeyFrame* k1 = as_keyframe(r1);
    KeyFrame* k2 = as_keyframe(r2);
    k1->mFeatVec = r1->F.mFeatVec; k2->mFeatVec = r2->F.mFeatVec;
    for (int i = 0; i < k2->N; i++)
        if (k2->mvpMapPoints[i]) k2->mvpMapPoints[i]->id = 2000000 + i;
    std::vector<MapPoint*> m12;
    Matcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(k1, k2, m12);
    for (int i = 0; i < k1->N; i++) out[i] = m12[i] ? m12[i]->id - 2000000 : -1;
    return n;
}


// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), pinhole keyframes
// (mpCamera2 == NULL).  matches12[i1] = i2 or -1 (vMatchedPairs = the pairs with i2 >= 0 in ascending i1).
// f12 / ep: the fundamental matrix Pinhole::epipolarConstrain builds and the epipole (:1063), for the caller's side.
int refm_search_for_triangulation(void* h1, void* h2, int onlyStereo, int coarse, float nnratio, int checkOri, int* matches12,
                                  float* f12, float* ep) {
    RefFrame *r1 = (RefFrame*)h1, *r2 = (RefFrame*)h2;
    KeyFrame* k1 = as_keyframe(r1);
    KeyFrame* k2 = as_keyframe(r2);
    k1->mFeatVec = r1->F.mFeatVec; k2->mFeatVec = r2->F.mFeatVec;
    k1->mpCamera2 = nullptr; k2->mpCamera2 = nullptr;
    std::vector<std::pair<size_t, size_t>> pairs;
    Matcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForTriangulation(k1, k2, pairs, onlyStereo != 0, coarse != 0);
    for (int i = 0; i < k1->N; i++) matches12[i] = -1;
    size_t prev = 0;
    for (size_t j = 0; j < pairs.size(); j++) {
        if (j && pairs[j].first <= prev) return -1000;   // must be ascending in i1
        prev = pairs[j].first;
        matches12[pairs[j].first] = (int)pairs[j].second;
    }
    // the same expressions as ORBmatcher.cc:1056-1075 and Pinhole.cpp:191-194, with the shim's matrix type
    Sophus::SE3f T12 = k1->GetPose() * k2->GetPoseInverse();
    Eigen::Matrix3f t12x = Sophus::SO3f::hat(T12.translation());
    Eigen::Matrix3f F12 = k1->mpCamera->toK_().transpose().inverse() * t12x * T12.rotationMatrix() * k2->mpCamera->toK_().inverse();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) f12[3 * i + j] = F12(i, j);
    Eigen::Vector2f e = k2->mpCamera->project(k2->GetPose() * k1->GetCameraCenter());
    ep[0] = e(0); ep[1] = e(1);
    return n;
}

}  // extern "C"
