// oracle/bow_oracle.h -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
//
// CPU restatement of the bag-of-words path: Frame::ComputeBoW / KeyFrame::ComputeBoW
// (/root/reference/src/Frame.cc:984-998, src/KeyFrame.cc:101-111) -> DBoW2::TemplatedVocabulary::transform
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1125-1197, per feature :1226-1258), FORB::distance
// (DBoW2/FORB.cpp:81-101), BowVector::addWeight / addIfNotExist / normalize (DBoW2/BowVector.cpp:34-84),
// FeatureVector::addFeature (DBoW2/FeatureVector.cpp:31-45), the text vocabulary format of
// loadFromTextFile (:1338-1424), and the two ORBmatcher::SearchByBoW overloads
// (src/ORBmatcher.cc:260-494, 893-1044).
// Parity status: PINNED against DBoW2 itself (every .cpp of Thirdparty/DBoW2 compiled verbatim into
// oracle/_ref/libref_orbmatcher.so) and the reference's own SearchByBoW bodies by
// tests/test_oracle_bow_vs_ref.py.
#pragma once
#include <cstdint>
#include <map>
#include <vector>

#include "match_oracle.h"

namespace bow_oracle {

// DBoW2 enums (DBoW2/BowVector.h:32-56)
enum { W_TF_IDF = 0, W_TF = 1, W_IDF = 2, W_BINARY = 3 };
enum { S_L1 = 0, S_L2 = 1, S_CHI = 2, S_KL = 3, S_BHATTA = 4, S_DOT = 5 };

struct Vocabulary {
    int k = 0, L = 0, scoring = S_L1, weighting = W_TF_IDF;
    std::vector<int> parent;                  // node 0 = root
    std::vector<std::vector<int>> children;   // in the order the nodes were added (ascending id)
    std::vector<uint8_t> desc;                // nNodes x 32
    std::vector<double> weight;
    std::vector<int> wordId;                  // -1 for inner nodes; leaves numbered in node-id order
    // nodes 1..n-1 from arrays indexed by node id (entry 0 = root, ignored)
    void build(int k_, int L_, int scoring_, int weighting_, int nNodes, const int* parent_, const uint8_t* desc_,
               const double* weight_);
    bool isLeaf(int id) const { return children[id].empty(); }
};

// transform(feature, word_id, weight, &nid, levelsup).  nid = -1 when the leaf is reached above the
// requested level (the reference leaves *nid uninitialised there).
void transform_feature(const Vocabulary& V, const uint8_t* d, int levelsup, int& word, double& w, int& nid);

// transform(features, BowVector&, FeatureVector&, levelsup)
void transform(const Vocabulary& V, const uint8_t* desc, int n, int levelsup, std::map<unsigned, double>& bow,
               std::map<unsigned, std::vector<unsigned>>& fv);

// Feature vectors in flat form: node ids ascending, start[i]..start[i+1] index into feat.
struct FeatVec {
    std::vector<int> node, start, feat;
};

struct BowSearchParams {
    int thLow;            // TH_LOW
    int strict;           // 1: bestDist1 <  TH_LOW (KeyFrame-KeyFrame, :973); 0: <= (KeyFrame-Frame, :351)
    float nnratio;
    int checkOrientation;
    int nLeftB;           // Frame::Nleft of the target (-1 = not fisheye); KeyFrame-Frame overload only
};
// Shared core of both SearchByBoW overloads.  A = the keyframe whose map points are looked up (validA[i] =
// it has a good map point), B = the frame / second keyframe (validB[i] = feature may be matched; all ones for
// the Frame overload).  matchA[iA] = matched B index or -1 (after the rotation histogram), matchAR[iA] = the
// right-camera match of the fisheye branch (:374-407) or -1.  Returns nmatches.
int search_by_bow(const FeatVec& fa, const uint8_t* descA, const float* angleA, const uint8_t* validA, int nA,
                  const FeatVec& fb, const uint8_t* descB, const float* angleB, const uint8_t* validB, int nB,
                  const BowSearchParams& prm, int* matchA, int* matchAR);

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1046-1324) for pinhole keyframes (mpCamera2 == NULL), with
// Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:186-216) given the fundamental matrix F12 (row major) it
// builds from K1, K2, R12, t12, and the epipole ep (:1063).  hasMp*[i] = the keyframe slot already holds a map point;
// uright* = mvuRight (NULL = monocular).  matches12[i1] = i2 or -1.  Returns nmatches.
struct TriSide {
    const FeatVec* fv;
    const match_oracle::OrbKp* keys;
    const uint8_t* desc;
    const float* uright;
    const uint8_t* hasMp;
    int n;
};
// Two-camera keyframes (mpCamera2 != NULL, :1071-1095, :1160-1241): keys / desc rows = [mvKeys | mvKeysRight]; per
// (bRight1, bRight2) combination k = 2 * bRight1 + bRight2 the KB8 parameters of the two cameras and R12 / t12
// (Tll, Tlr, Trl, Trr); the epipolar gate is KannalaBrandt8::epipolarConstrain (KannalaBrandt8.cpp:322-328).
struct TriRig {
    int nLeft1, nLeft2;
    const float* levelSigma2A;
    float P1[4][8], P2[4][8], prec1[4], prec2[4], R12[4][9], t12[4][3];
};
int search_for_triangulation(const TriSide& A, const TriSide& B, const float* F12, const float* ep, const float* scaleFactorsB,
                             const float* levelSigma2B, int onlyStereo, int coarse, int checkOrientation, int thLow,
                             int* matches12, const TriRig* rig = nullptr);

}  // namespace bow_oracle
