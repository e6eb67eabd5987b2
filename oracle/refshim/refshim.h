// oracle/refshim/refshim.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-ins for the classes the reference's matcher functions touch (Frame, MapPoint, KeyFrame,
// Eigen::Vector{2,3}f, Sophus::SE3f, GeometricCamera), so that the reference's OWN function bodies --
// sliced at build time out of /root/reference/src/{ORBmatcher,Frame,MapPoint}.cc by oracle/ref_slices.py,
// never copied into this repo -- compile without Eigen/Sophus/DBoW2/g2o, which the image does not have.
// Only the data members and accessors those bodies read are declared (names as in the reference's
// include/Frame.h, include/MapPoint.h, include/KeyFrame.h); everything else of the reference is absent.
// Poses are translation-only and the camera is a plain pinhole: the tests pin the MATCHING logic
// (windows, level filters, claims, ratio tests, rotation histogram, stereo partners), not the geometry
// the caller keeps on the host.
#pragma once
// The reference's ORBmatcher.h pulls "MapPoint.h", "KeyFrame.h" and "Frame.h" with quote includes, which
// resolve next to it before any -I path: pre-define their include guards so the real ones are no-ops.
#define MAPPOINT_H
#define KEYFRAME_H
#define FRAME_H
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

#include "ORBextractor.h"
// DBoW2's own headers (the reference vendors DBoW2; Boost.Serialization is stubbed in refshim/boost)
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;  // the reference's headers rely on this leaking from its own includes

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {
struct Vector3f {
    float v[3];
    Vector3f() : v{0, 0, 0} {}
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3f operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    float dot(const Vector3f& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    Vector3f operator/(float s) const { return Vector3f(v[0] / s, v[1] / s, v[2] / s); }
    Vector3f operator*(float s) const { return Vector3f(v[0] * s, v[1] * s, v[2] * s); }
};
struct Matrix3f {
    float m[3][3];
    Matrix3f() : m{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}} {}
    float& operator()(int i, int j) { return m[i][j]; }
    float operator()(int i, int j) const { return m[i][j]; }
    Matrix3f transpose() const {
        Matrix3f r;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[j][i];
        return r;
    }
    Matrix3f operator*(const Matrix3f& o) const {
        Matrix3f r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j];
        return r;
    }
    Vector3f operator*(const Vector3f& v) const {
        return Vector3f(m[0][0] * v(0) + m[0][1] * v(1) + m[0][2] * v(2), m[1][0] * v(0) + m[1][1] * v(1) + m[1][2] * v(2),
                        m[2][0] * v(0) + m[2][1] * v(1) + m[2][2] * v(2));
    }
    Matrix3f inverse() const {   // adjugate / determinant (the tests only need SOME fixed F12, shared by both sides)
        Matrix3f r;
        const float d = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                        m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
        r.m[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / d; r.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / d;
        r.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / d; r.m[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / d;
        r.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / d; r.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / d;
        r.m[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / d; r.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / d;
        r.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / d;
        return r;
    }
};
struct Vector2f {
    float v[2];
    Vector2f() : v{0, 0} {}
    Vector2f(float x, float y) : v{x, y} {}
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};
}  // namespace Eigen

namespace Sophus {
// translation-only rigid transform
template <class T>
struct SE3 {
    Eigen::Vector3f t;
    SE3() {}
    explicit SE3(const Eigen::Vector3f& t_) : t(t_) {}
    SE3(const Eigen::Matrix3f&, const Eigen::Vector3f& t_) : t(t_) {}
    SE3 inverse() const { return SE3(Eigen::Vector3f(-t(0), -t(1), -t(2))); }
    SE3 operator*(const SE3& o) const { return SE3(o.t + t); }
    Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
    Eigen::Vector3f translation() const { return t; }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return p + t; }
};
typedef SE3<float> SE3f;
struct SO3f {
    static Eigen::Matrix3f hat(const Eigen::Vector3f& v) {
        Eigen::Matrix3f r;
        r(0, 0) = 0; r(0, 1) = -v(2); r(0, 2) = v(1);
        r(1, 0) = v(2); r(1, 1) = 0; r(1, 2) = -v(0);
        r(2, 0) = -v(1); r(2, 1) = v(0); r(2, 2) = 0;
        return r;
    }
};
// scale + translation similarity: p -> s*p + t
template <class T>
struct Sim3 {
    float s = 1;
    Eigen::Vector3f t;
    Sim3() {}
    Sim3(float s_, const Eigen::Vector3f& t_) : s(s_), t(t_) {}
    Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
    Eigen::Vector3f translation() const { return t; }
    float scale() const { return s; }
    Sim3 inverse() const { return Sim3(1.0f / s, Eigen::Vector3f(-t(0) / s, -t(1) / s, -t(2) / s)); }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return p * s + t; }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

#include <tuple>

namespace ORB_SLAM3 {

class Frame;
class KeyFrame;
class MapPoint;

struct RefAction { int kind, a, b; };   // 1: a->Replace(b)   2: a->AddObservation(kf, idx=b)   3: kf->AddMapPoint(a, idx=b)
extern std::vector<RefAction> g_refActions;

class GeometricCamera {
   public:
    float fx = 1, fy = 1, cx = 0, cy = 0;
    virtual ~GeometricCamera() {}
    Eigen::Vector2f project(const Eigen::Vector3f& p) const {
        return Eigen::Vector2f(fx * p(0) / p(2) + cx, fy * p(1) / p(2) + cy);
    }
    virtual Eigen::Matrix3f toK_() = 0;
    virtual bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2,
                                   const Eigen::Matrix3f& R12, const Eigen::Vector3f& t12, const float sigmaLevel,
                                   const float unc) = 0;
};
class Pinhole : public GeometricCamera {
   public:
    Eigen::Matrix3f toK_() override {   // src/CameraModels/Pinhole.cpp:168-173
        Eigen::Matrix3f K;
        K(0, 0) = fx; K(0, 1) = 0; K(0, 2) = cx; K(1, 0) = 0; K(1, 1) = fy; K(1, 2) = cy; K(2, 0) = 0; K(2, 1) = 0; K(2, 2) = 1;
        return K;
    }
    // body sliced from src/CameraModels/Pinhole.cpp
    bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2,
                           const Eigen::Matrix3f& R12, const Eigen::Vector3f& t12, const float sigmaLevel,
                           const float unc) override;
};

class MapPoint {
   public:
    // members the matchers read directly (include/MapPoint.h)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 1, mTrackViewCosR = 1;
    // accessors
    bool isBad() { return bad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return desc.clone(); }
    Eigen::Vector3f GetWorldPos() { return pos; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float& currentDist, Frame* pF);      // bodies sliced from src/MapPoint.cc
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    Eigen::Vector3f GetNormal() { return normal; }
    bool IsInKeyFrame(KeyFrame* pKF) { return inKF; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF) { return std::tuple<int, int>(idxInOtherKF, -1); }
    // graph updates of Fuse are recorded, not performed: (kind, this->id, other id or keypoint index)
    void Replace(MapPoint* pMP);
    void AddObservation(KeyFrame* pKF, int idx);
    void ComputeDistinctiveDescriptors();                       // body sliced from src/MapPoint.cc
    std::map<KeyFrame*, std::tuple<int, int>> mObservations;
    std::mutex mMutexFeatures;
    bool mbBad = false;
    cv::Mat mDescriptor;
    int id = -1;
    bool inKF = false;
    int idxInOtherKF = -1;
    Eigen::Vector3f normal;
    // state
    std::mutex mMutexPos;
    float mfMinDistance = 0, mfMaxDistance = 0;
    bool bad = false;
    int nObs = 1;
    cv::Mat desc;
    Eigen::Vector3f pos;
};

class KeyFrame {
   public:
    // sliced from src/KeyFrame.cc
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool IsInImage(const float& x, const float& y) const;

    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() {   // src/KeyFrame.cc:370-385
        std::set<MapPoint*> s;
        for (MapPoint* p : mvpMapPoints)
            if (p && !p->isBad()) s.insert(p);
        return s;
    }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    bool isBad() { return mbBadKF; }
    bool mbBadKF = false;
    void AddMapPoint(MapPoint* pMP, const size_t& idx) {
        mvpMapPoints[idx] = pMP;
        g_refActions.push_back({3, pMP->id, (int)idx});
    }
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    Sophus::SE3f GetRightPose() { return mTcw; }
    Sophus::SE3f GetRightPoseInverse() { return mTcw.inverse(); }
    Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }
    Eigen::Vector3f GetRightCameraCenter() { return mTcw.inverse().translation(); }

    float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    int N = 0, NLeft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvInvLevelSigma2, mvLevelSigma2;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;      // ints in the reference's KeyFrame (include/KeyFrame.h:403)
    int mnGridCols = 64, mnGridRows = 48;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::vector<std::vector<size_t>>> mGrid, mGridRight;
    std::vector<MapPoint*> mvpMapPoints;
    Sophus::SE3f mTcw;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
};

inline void MapPoint::Replace(MapPoint* pMP) { g_refActions.push_back({1, id, pMP->id}); }
inline void MapPoint::AddObservation(KeyFrame*, int idx) { g_refActions.push_back({2, id, idx}); }

class Frame {
   public:
    // sliced from src/Frame.cc
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                     const int maxLevel = -1, const bool bRight = false) const;
    void ComputeStereoMatches();
    void AssignFeaturesToGrid();

    Sophus::SE3f GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }

    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    GeometricCamera* mpCamera = nullptr;
    float mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    int mnScaleLevels = 0;
    float mfScaleFactor = 0, mfLogScaleFactor = 0;
    vector<float> mvScaleFactors, mvInvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    int Nleft = -1, Nright = -1;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    Sophus::SE3f mTcw, mTrl;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    GeometricCamera* mpCamera2 = nullptr;
};

}  // namespace ORB_SLAM3
