// oracle/bow_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  See bow_oracle.h.
#include "bow_oracle.h"

#include <cmath>

namespace bow_oracle {

using match_oracle::descriptor_distance;

void Vocabulary::build(int k_, int L_, int scoring_, int weighting_, int nNodes, const int* parent_,
                       const uint8_t* desc_, const double* weight_) {
    k = k_; L = L_; scoring = scoring_; weighting = weighting_;
    parent.assign(parent_, parent_ + nNodes);
    desc.assign(desc_, desc_ + (size_t)nNodes * 32);
    weight.assign(weight_, weight_ + nNodes);
    children.assign(nNodes, {});
    for (int i = 1; i < nNodes; i++) children[parent[i]].push_back(i);  // loadFromTextFile :1390
    wordId.assign(nNodes, -1);
    int w = 0;
    for (int i = 1; i < nNodes; i++)
        if (children[i].empty()) wordId[i] = w++;  // :1409-1416 (leaf flag of the file == no children)
}

// TemplatedVocabulary.h:1226-1258
void transform_feature(const Vocabulary& V, const uint8_t* d, int levelsup, int& word, double& w, int& nid) {
    const int nid_level = V.L - levelsup;
    nid = -1;
    if (nid_level <= 0) nid = 0;
    int final_id = 0, current_level = 0;
    do {
        ++current_level;
        const std::vector<int>& nodes = V.children[final_id];
        final_id = nodes[0];
        double best_d = descriptor_distance(d, &V.desc[32 * (size_t)final_id]);
        for (size_t c = 1; c < nodes.size(); c++) {
            const int id = nodes[c];
            const double dd = descriptor_distance(d, &V.desc[32 * (size_t)id]);
            if (dd < best_d) { best_d = dd; final_id = id; }
        }
        if (current_level == nid_level) nid = final_id;
    } while (!V.isLeaf(final_id));
    word = V.wordId[final_id];
    w = V.weight[final_id];
}

// TemplatedVocabulary.h:1125-1197
void transform(const Vocabulary& V, const uint8_t* desc, int n, int levelsup, std::map<unsigned, double>& bow,
               std::map<unsigned, std::vector<unsigned>>& fv) {
    bow.clear();
    fv.clear();
    if (V.children.empty() || V.children[0].empty()) return;
    const bool must = V.scoring != S_DOT;               // ScoringObject.h: every scoring but DotProduct normalises
    const bool l1 = V.scoring != S_L2;
    const bool tf = V.weighting == W_TF || V.weighting == W_TF_IDF;
    for (int i = 0; i < n; i++) {
        int id, nid;
        double w;
        transform_feature(V, desc + 32 * (size_t)i, levelsup, id, w, nid);
        if (w > 0) {
            if (tf) bow[(unsigned)id] += w;              // addWeight, BowVector.cpp:34-46
            else bow.insert({(unsigned)id, w});          // addIfNotExist, :50-58
            fv[(unsigned)nid].push_back((unsigned)i);    // addFeature, FeatureVector.cpp:31-45
        }
    }
    if (tf && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (auto& e : bow) e.second /= nd;
    }
    if (must) {                                           // BowVector::normalize, BowVector.cpp:62-84
        double norm = 0.0;
        if (l1) for (auto& e : bow) norm += std::fabs(e.second);
        else { for (auto& e : bow) norm += e.second * e.second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (auto& e : bow) e.second /= norm;
    }
}

static void three_maxima(const int* cnt, int Lh, int& ind1, int& ind2, int& ind3) {  // ORBmatcher.cc:2336-2378
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < Lh; i++) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// ORBmatcher.cc:260-494 (A = pKF, B = F) and :893-1044 (A = pKF1, B = pKF2)
int search_by_bow(const FeatVec& fa, const uint8_t* descA, const float* angleA, const uint8_t* validA, int nA,
                  const FeatVec& fb, const uint8_t* descB, const float* angleB, const uint8_t* validB, int nB,
                  const BowSearchParams& prm, int* matchA, int* matchAR) {
    enum { HISTO = 30 };
    for (int i = 0; i < nA; i++) { matchA[i] = -1; if (matchAR) matchAR[i] = -1; }
    std::vector<uint8_t> taken(nB, 0);
    std::vector<int> rotHist[HISTO];   // holds (iA << 1 | isRight)
    const float factor = 1.0f / HISTO;
    int nmatches = 0;
    size_t ia = 0, ib = 0;
    const size_t na = fa.node.size(), nb = fb.node.size();
    auto vote = [&](int iA, int iB, int right) {
        float rot = angleA[iA] - angleB[iB];
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)roundf(rot * factor);
        if (bin == HISTO) bin = 0;
        rotHist[bin].push_back(iA * 2 + right);
    };
    while (ia < na && ib < nb) {
        if (fa.node[ia] == fb.node[ib]) {
            for (int pa = fa.start[ia]; pa < fa.start[ia + 1]; pa++) {
                const int iA = fa.feat[pa];
                if (!validA[iA]) continue;
                const uint8_t* dA = descA + 32 * (size_t)iA;
                int bestDist1 = 256, bestIdx = -1, bestDist2 = 256, bestDist1R = 256, bestIdxR = -1, bestDist2R = 256;
                for (int pb = fb.start[ib]; pb < fb.start[ib + 1]; pb++) {
                    const int iB = fb.feat[pb];
                    if (taken[iB] || (validB && !validB[iB])) continue;
                    const int dist = descriptor_distance(dA, descB + 32 * (size_t)iB);
                    if (prm.nLeftB == -1 || iB < prm.nLeftB) {
                        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx = iB; }
                        else if (dist < bestDist2) bestDist2 = dist;
                    } else {
                        if (dist < bestDist1R) { bestDist2R = bestDist1R; bestDist1R = dist; bestIdxR = iB; }
                        else if (dist < bestDist2R) bestDist2R = dist;
                    }
                }
                const bool pass = prm.strict ? bestDist1 < prm.thLow : bestDist1 <= prm.thLow;
                if (pass) {
                    if ((float)bestDist1 < prm.nnratio * (float)bestDist2) {
                        matchA[iA] = bestIdx;
                        taken[bestIdx] = 1;
                        if (prm.checkOrientation) vote(iA, bestIdx, 0);
                        nmatches++;
                    }
                    if (prm.nLeftB != -1 && bestDist1R <= prm.thLow) {   // :374-407, ratio test disabled by `|| true`
                        matchAR[iA] = bestIdxR;
                        taken[bestIdxR] = 1;
                        if (prm.checkOrientation) vote(iA, bestIdxR, 1);
                        nmatches++;
                    }
                }
            }
            ia++; ib++;
        } else if (fa.node[ia] < fb.node[ib]) {
            while (ia < na && fa.node[ia] < fb.node[ib]) ia++;   // lower_bound
        } else {
            while (ib < nb && fb.node[ib] < fa.node[ia]) ib++;
        }
    }
    if (prm.checkOrientation) {
        int cnt[HISTO], ind1, ind2, ind3;
        for (int i = 0; i < HISTO; i++) cnt[i] = (int)rotHist[i].size();
        three_maxima(cnt, HISTO, ind1, ind2, ind3);
        for (int i = 0; i < HISTO; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int e : rotHist[i]) {
                if (e & 1) matchAR[e >> 1] = -1; else matchA[e >> 1] = -1;
                nmatches--;
            }
        }
    }
    return nmatches;
}

}  // namespace bow_oracle
namespace kb8_oracle {   // kb8_oracle.cpp
float triangulate_matches(const float* P1, float prec1, const float* P2, float prec2, const float* R12, const float* t12,
                          const float* pt1, const float* pt2, float sigmaLevel, float unc, float* p3D);
}
namespace bow_oracle {

// ORBmatcher.cc:1046-1324; Pinhole.cpp:196-215 for mpCamera2 == NULL, KannalaBrandt8.cpp:322-328 for two-camera keyframes
int search_for_triangulation(const TriSide& A, const TriSide& B, const float* F12, const float* ep, const float* scaleFactorsB,
                             const float* levelSigma2B, int onlyStereo, int coarse, int checkOrientation, int thLow,
                             int* matches12, const TriRig* rig) {
    enum { HISTO = 30 };
    for (int i = 0; i < A.n; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO];
    const float factor = 1.0f / HISTO;
    int nmatches = 0;
    const FeatVec &fa = *A.fv, &fb = *B.fv;
    size_t ia = 0, ib = 0;
    const size_t na = fa.node.size(), nb = fb.node.size();
    while (ia < na && ib < nb) {
        if (fa.node[ia] == fb.node[ib]) {
            for (int pa = fa.start[ia]; pa < fa.start[ia + 1]; pa++) {
                const int idx1 = fa.feat[pa];
                if (A.hasMp[idx1]) continue;
                const bool bStereo1 = !rig && A.uright && A.uright[idx1] >= 0;   // :1121 (!pKF1->mpCamera2 && mvuRight >= 0)
                if (onlyStereo && !bStereo1) continue;
                const bool bRight1 = rig && idx1 >= rig->nLeft1;                 // :1126-1128
                const match_oracle::OrbKp& kp1 = A.keys[idx1];
                const uint8_t* d1 = A.desc + 32 * (size_t)idx1;
                int bestDist = thLow, bestIdx2 = -1;
                for (int pb = fb.start[ib]; pb < fb.start[ib + 1]; pb++) {
                    const int idx2 = fb.feat[pb];
                    if (B.hasMp[idx2]) continue;                 // vbMatched2 is never set in this version of the loop
                    const bool bStereo2 = !rig && B.uright && B.uright[idx2] >= 0;
                    if (onlyStereo && !bStereo2) continue;
                    const int dist = descriptor_distance(d1, B.desc + 32 * (size_t)idx2);
                    if (dist > thLow || dist > bestDist) continue;
                    const match_oracle::OrbKp& kp2 = B.keys[idx2];
                    if (!bStereo1 && !bStereo2 && !rig) {                        // :1196
                        const float distex = ep[0] - kp2.x;
                        const float distey = ep[1] - kp2.y;
                        if (distex * distex + distey * distey < 100 * scaleFactorsB[kp2.octave]) continue;
                    }
                    bool ok = coarse != 0;
                    if (!ok && rig) {   // :1205-1241 camera / pose selection, then KannalaBrandt8::epipolarConstrain
                        const bool bRight2 = idx2 >= rig->nLeft2;
                        const int k = 2 * (bRight1 ? 1 : 0) + (bRight2 ? 1 : 0);
                        const float pt1[2] = {kp1.x, kp1.y}, pt2[2] = {kp2.x, kp2.y};
                        float p3D[3];
                        ok = kb8_oracle::triangulate_matches(rig->P1[k], rig->prec1[k], rig->P2[k], rig->prec2[k], rig->R12[k], rig->t12[k],
                                                             pt1, pt2, rig->levelSigma2A[kp1.octave], levelSigma2B[kp2.octave], p3D) > 0.0001f;
                    } else if (!ok) {   // Pinhole::epipolarConstrain
                        const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                        const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                        const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                        const float num = a * kp2.x + b * kp2.y + c;
                        const float den = a * a + b * b;
                        if (den == 0) ok = false;
                        else {
                            const float dsqr = num * num / den;
                            ok = dsqr < 3.84 * levelSigma2B[kp2.octave];
                        }
                    }
                    if (ok) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2;
                    nmatches++;
                    if (checkOrientation) {
                        float rot = kp1.angle - B.keys[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)roundf(rot * factor);
                        if (bin == HISTO) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            ia++; ib++;
        } else if (fa.node[ia] < fb.node[ib]) {
            while (ia < na && fa.node[ia] < fb.node[ib]) ia++;
        } else {
            while (ib < nb && fb.node[ib] < fa.node[ia]) ib++;
        }
    }
    if (checkOrientation) {
        int cnt[HISTO], ind1, ind2, ind3;
        for (int i = 0; i < HISTO; i++) cnt[i] = (int)rotHist[i].size();
        three_maxima(cnt, HISTO, ind1, ind2, ind3);
        for (int i = 0; i < HISTO; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // namespace bow_oracle
