// oracle/match_oracle.cpp -- TEST INFRASTRUCTURE ONLY. See match_oracle.h.
#include "match_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <utility>

#include "cvprims.h"

namespace match_oracle {

// reference src/ORBmatcher.cc:2384-2404 (bit-hack popcount over 8 x 32-bit words)
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4);
        memcpy(&pb, b + 4 * i, 4);
        unsigned int v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// reference src/Frame.cc:469-504 with PosInGrid :962-978
void FrameView::assign_features_to_grid() {
    for (int i = 0; i < GRID_COLS; i++)
        for (int j = 0; j < GRID_ROWS; j++) grid[i][j].clear();
    for (int i = 0; i < N; i++) {
        const OrbKp& kp = keys[i];
        int posX = (int)roundf((kp.x - minX) * gridWInv);
        int posY = (int)roundf((kp.y - minY) * gridHInv);
        if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
        grid[posX][posY].push_back(i);
    }
}

// reference src/Frame.cc:859-951
std::vector<int> FrameView::features_in_area(float x, float y, float r, int minLevel,
                                             int maxLevel) const {
    std::vector<int> vIndices;
    float factorX = r, factorY = r;
    const int nMinCellX = std::max(0, (int)floorf((x - minX - factorX) * gridWInv));
    if (nMinCellX >= GRID_COLS) return vIndices;
    const int nMaxCellX = std::min((int)GRID_COLS - 1, (int)ceilf((x - minX + factorX) * gridWInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floorf((y - minY - factorY) * gridHInv));
    if (nMinCellY >= GRID_ROWS) return vIndices;
    const int nMaxCellY = std::min((int)GRID_ROWS - 1, (int)ceilf((y - minY + factorY) * gridHInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<int>& vCell = grid[ix][iy];
            for (size_t j = 0; j < vCell.size(); j++) {
                const OrbKp& kpUn = keys[vCell[j]];
                if (bCheckLevels) {
                    if (kpUn.octave < minLevel) continue;
                    if (maxLevel >= 0)
                        if (kpUn.octave > maxLevel) continue;
                }
                const float distx = kpUn.x - x;
                const float disty = kpUn.y - y;
                if (fabsf(distx) < factorX && fabsf(disty) < factorY) vIndices.push_back(vCell[j]);
            }
        }
    return vIndices;
}

// reference src/ORBmatcher.cc:2336-2378
static void compute_three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2,
                                 int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) {
            max3 = max2; max2 = max1; max1 = s;
            ind3 = ind2; ind2 = ind1; ind1 = i;
        } else if (s > max2) {
            max3 = max2; max2 = s;
            ind3 = ind2; ind2 = i;
        } else if (s > max3) {
            max3 = s; ind3 = i;
        }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// reference src/ORBmatcher.cc:46-170 (mode 0), :1951-2088 + :2157-2185 (mode 1),
// :2197-2325 (mode 2); Nleft == -1 branches.
int search_by_projection(FrameView& F, const std::vector<ProjPoint>& pts, const uint8_t* pdesc,
                         const SearchParams& prm, const uint8_t* claimed0, int* assigned,
                         int* best_idx, int* best_dist) {
    int nmatches = 0;
    std::vector<uint8_t> claimed(claimed0, claimed0 + F.N);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    for (size_t j = 0; j < pts.size(); j++) {
        const ProjPoint& p = pts[j];
        best_idx[j] = -1;
        best_dist[j] = 256;
        if (!p.valid) continue;
        const std::vector<int> vIndices = F.features_in_area(p.u, p.v, p.radius, p.minLevel, p.maxLevel);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = pdesc + 32 * j;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (claimed[idx]) continue;
            if (prm.mode != 2 && F.uright && F.uright[idx] > 0) {
                const float er = fabsf(p.ur - F.uright[idx]);
                if (er > p.radius) continue;
            }
            const int dist = descriptor_distance(dMP, F.desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F.keys[idx].octave;
                bestIdx = idx;
            } else if (prm.mode == 0 && dist < bestDist2) {
                bestLevel2 = F.keys[idx].octave;
                bestDist2 = dist;
            }
        }
        best_dist[j] = bestDist;
        if (bestDist <= prm.thAccept) {
            if (prm.mode == 0) {
                if (bestLevel == bestLevel2 && bestDist > prm.nnratio * bestDist2) continue;
            }
            assigned[bestIdx] = (int)j;
            claimed[bestIdx] = prm.mode == 2 ? 1 : p.blocks;
            best_idx[j] = bestIdx;
            nmatches++;
            if (prm.mode != 0 && prm.checkOrientation) {
                float rot = p.angle - F.keys[bestIdx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx);
            }
        }
    }
    if (prm.mode != 0 && prm.checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int k : rotHist[i]) {
                    assigned[k] = -1;
                    nmatches--;
                }
    }
    return nmatches;
}

// reference src/ORBmatcher.cc:46-240 (mode 0) and :1951-2185 (mode 1) with F.Nleft != -1
int search_by_projection_fisheye(FrameView& FL, FrameView& FR, const int* l2r, const int* r2l,
                                 const std::vector<ProjPoint>& ptsL, const std::vector<ProjPoint>& ptsR,
                                 const uint8_t* pdesc, const SearchParams& prm, const uint8_t* claimed0,
                                 int* assigned, int* best_idx_l, int* best_idx_r) {
    const int Nl = FL.N, N = FL.N + FR.N;
    int nmatches = 0;
    std::vector<uint8_t> claimed(claimed0, claimed0 + N);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    // one window search; slotOff maps a keypoint index of `F` to its F.mvpMapPoints slot
    auto window = [&](FrameView& F, int slotOff, const ProjPoint& p, const uint8_t* dMP, bool second, bool& had,
                      int& bestDist, int& bestLevel, int& bestDist2, int& bestLevel2, int& bestIdx) {
        bestDist = 256; bestLevel = -1; bestDist2 = 256; bestLevel2 = -1; bestIdx = -1;
        const std::vector<int> vIndices = F.features_in_area(p.u, p.v, p.radius, p.minLevel, p.maxLevel);
        had = !vIndices.empty();
        for (int idx : vIndices) {
            if (claimed[idx + slotOff]) continue;
            const int dist = descriptor_distance(dMP, F.desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F.keys[idx].octave;
                bestIdx = idx;
            } else if (second && dist < bestDist2) {
                bestLevel2 = F.keys[idx].octave;
                bestDist2 = dist;
            }
        }
    };
    auto vote = [&](float angLast, float angCur, int slot) {
        float rot = angLast - angCur;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)roundf(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(slot);
    };
    for (size_t j = 0; j < ptsL.size(); j++) {
        const ProjPoint &pL = ptsL[j], &pR = ptsR[j];
        const uint8_t* dMP = pdesc + 32 * j;
        best_idx_l[j] = -1;
        best_idx_r[j] = -1;
        int bD, bL, bD2, bL2, bI;
        bool had = false;
        if (prm.mode == 0) {
            if (pL.valid) {
                window(FL, 0, pL, dMP, true, had, bD, bL, bD2, bL2, bI);
                if (had && bD <= prm.thAccept) {
                    if (bL == bL2 && bD > prm.nnratio * bD2) continue;   // :154 skips the right camera too
                    assigned[bI] = (int)j; claimed[bI] = pL.blocks;
                    if (l2r[bI] != -1) { assigned[l2r[bI] + Nl] = (int)j; claimed[l2r[bI] + Nl] = pL.blocks; nmatches++; }
                    nmatches++;
                    best_idx_l[j] = bI;
                }
            }
            if (pR.valid) {
                window(FR, Nl, pR, dMP, true, had, bD, bL, bD2, bL2, bI);
                if (!had) continue;
                if (bD <= prm.thAccept) {
                    if (bL == bL2 && bD > prm.nnratio * bD2) continue;
                    if (r2l[bI] != -1) { assigned[r2l[bI]] = (int)j; claimed[r2l[bI]] = pL.blocks; nmatches++; }
                    assigned[bI + Nl] = (int)j; claimed[bI + Nl] = pL.blocks;
                    nmatches++;
                    best_idx_r[j] = bI;
                }
            }
        } else {
            if (!pL.valid) continue;
            window(FL, 0, pL, dMP, false, had, bD, bL, bD2, bL2, bI);
            if (!had) continue;                                          // :2027 skips the right camera too
            if (bD <= prm.thAccept) {
                assigned[bI] = (int)j; claimed[bI] = pL.blocks;
                nmatches++;
                best_idx_l[j] = bI;
                if (prm.checkOrientation) vote(pL.angle, FL.keys[bI].angle, bI);
            }
            window(FR, Nl, pR, dMP, false, had, bD, bL, bD2, bL2, bI);
            if (bD <= prm.thAccept) {
                assigned[bI + Nl] = (int)j; claimed[bI + Nl] = pL.blocks;
                nmatches++;
                best_idx_r[j] = bI;
                if (prm.checkOrientation) vote(pL.angle, FR.keys[bI].angle, bI + Nl);
            }
        }
    }
    if (prm.mode != 0 && prm.checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int k : rotHist[i]) {
                    assigned[k] = -1;
                    nmatches--;
                }
    }
    return nmatches;
}

// ORBmatcher.cc:1409-1478 (Fuse), :1592-1616 (Fuse, Sim3), :1755-1780 / :1853-1878 (SearchBySim3),
// :552-581 (SearchByProjection, Sim3; without its vpMatched skip, which is search_by_projection mode 2).
void search_window(const FrameView& KF, const std::vector<ProjPoint>& pts, const uint8_t* pdesc,
                   const WindowParams& prm, int* best_idx, int* best_dist) {
    for (size_t j = 0; j < pts.size(); j++) {
        const ProjPoint& p = pts[j];
        best_idx[j] = -1;
        best_dist[j] = 256;
        if (!p.valid) continue;
        // KeyFrame::GetFeaturesInArea has no level arguments (KeyFrame.cc:843): all levels, then the filter
        const std::vector<int> vIndices = KF.features_in_area(p.u, p.v, p.radius, -1, -1);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = pdesc + 32 * j;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            const OrbKp& kp = KF.keys[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < p.minLevel || kpLevel > p.maxLevel) continue;
            if (prm.fuseGate) {
                const float ex = p.u - kp.x;
                const float ey = p.v - kp.y;
                if (KF.uright && KF.uright[idx] >= 0) {
                    const float er = p.ur - KF.uright[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * prm.invLevelSigma2[kpLevel] > 7.8) continue;
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * prm.invLevelSigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(dMP, KF.desc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        best_dist[j] = bestDist;
        if (bestDist <= prm.thAccept) best_idx[j] = bestIdx;
    }
}

int search_by_sim3(const FrameView& KF1, const FrameView& KF2, const std::vector<ProjPoint>& pts12,
                   const uint8_t* desc1, const std::vector<ProjPoint>& pts21, const uint8_t* desc2,
                   int thAccept, int* match12) {
    const int N1 = (int)pts12.size(), N2 = (int)pts21.size();
    std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1), d1(N1), d2(N2);
    const WindowParams wp{thAccept, 0, nullptr, 0};
    search_window(KF2, pts12, desc1, wp, vnMatch1.data(), d1.data());  // :1729-1812
    search_window(KF1, pts21, desc2, wp, vnMatch2.data(), d2.data());  // :1826-1909
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {  // :1922-1937
        match12[i1] = -1;
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0) {
            const int idx1 = vnMatch2[idx2];
            if (idx1 == i1) { match12[i1] = idx2; nFound++; }
        }
    }
    return nFound;
}

// reference src/MapPoint.cc:476-521
int distinctive_descriptor(const uint8_t* desc, int N) {
    if (N <= 0) return -1;
    std::vector<float> Distances((size_t)N * N);
    for (int i = 0; i < N; i++) {
        Distances[(size_t)i * N + i] = 0;
        for (int j = i + 1; j < N; j++) {
            const int distij = descriptor_distance(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
            Distances[(size_t)i * N + j] = (float)distij;
            Distances[(size_t)j * N + i] = (float)distij;
        }
    }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; i++) {
        std::vector<int> vDists(Distances.begin() + (size_t)i * N, Distances.begin() + (size_t)(i + 1) * N);
        std::sort(vDists.begin(), vDists.end());
        const int median = vDists[(size_t)(0.5 * (N - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    return BestIdx;
}

// reference src/ORBmatcher.cc:735-891
int search_for_initialization(const FrameView& F1, FrameView& F2, float* vbPrevMatched, int windowSize,
                              float mfNNratio, bool mbCheckOrientation, int* vnMatches12) {
    int nmatches = 0;
    for (int i = 0; i < F1.N; i++) vnMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(F2.N, INT_MAX);
    std::vector<int> vnMatches21(F2.N, -1);
    for (int i1 = 0; i1 < F1.N; i1++) {
        const OrbKp& kp1 = F1.keys[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        const std::vector<int> vIndices2 = F2.features_in_area(vbPrevMatched[2 * i1], vbPrevMatched[2 * i1 + 1],
                                                               (float)windowSize, level1, level1);
        if (vIndices2.empty()) continue;
        const uint8_t* d1 = F1.desc + 32 * (size_t)i1;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            const int dist = descriptor_distance(d1, F2.desc + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * mfNNratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (mbCheckOrientation) {
                    float rot = F1.keys[i1].angle - F2.keys[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (mbCheckOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < F1.N; i1++)
        if (vnMatches12[i1] >= 0) {
            vbPrevMatched[2 * i1] = F2.keys[vnMatches12[i1]].x;
            vbPrevMatched[2 * i1 + 1] = F2.keys[vnMatches12[i1]].y;
        }
    return nmatches;
}

// reference src/Frame.cc:1102-1358
void compute_stereo_matches(const OrbKp* keysL, const uint8_t* descL, int N, const OrbKp* keysR,
                            const uint8_t* descR, int Nr, const PyrLevelView* pyrL,
                            const PyrLevelView* pyrR, const float* mvScaleFactors,
                            const float* mvInvScaleFactors, float mbf, float mb, float* mvuRight,
                            float* mvDepth) {
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = pyrL[0].h;
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const OrbKp& kp = keysR[iR];
        const float kpY = kp.y;
        const float r = 2.0f * mvScaleFactors[kp.octave];
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows)  // the reference indexes unchecked (UB outside)
                vRowIndices[yi].push_back(iR);
    }
    const float minZ = mb;
    const float minD = 0;
    const float maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const OrbKp& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y;
        const float uL = kpL.x;
        if ((int)vL < 0 || (int)vL >= nRows) continue;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const uint8_t* dL = descL + 32 * (size_t)iL;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const OrbKp& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, descR + 32 * iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = keysR[bestIdxR].x;
            const float scaleFactor = mvInvScaleFactors[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const PyrLevelView& PL = pyrL[kpL.octave];
            const PyrLevelView& PR = pyrR[kpL.octave];
            int bestDistS = INT_MAX;
            int bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= PR.w) continue;
            const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w);
            for (int incR = -L; incR <= +L; incR++) {
                const int xr0 = (int)(scaleduR0 + incR - w);
                int sad = 0;  // cv::norm(IL, IR, NORM_L1) on CV_8U
                for (int yy = 0; yy < 2 * w + 1; yy++)
                    for (int xx = 0; xx < 2 * w + 1; xx++)
                        sad += std::abs((int)PL.roi[(y0 + yy) * PL.step + xl0 + xx] -
                                        (int)PR.roi[(y0 + yy) * PR.step + xr0 + xx]);
                float dist = (float)(double)sad;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = mvScaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = uL - 0.01;
                }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return;  // the reference reads vDistIdx[0] here (upstream bug)
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
    }
}

// reference src/Frame.cc:1553-1562
void fisheye_ratio_matches(const uint8_t* q, int nq, const uint8_t* t, int nt, int* match,
                           int* idx2, int* dist2) {
    cvp::bf_knn2(q, nq, t, nt, idx2, dist2);
    for (int i = 0; i < nq; i++) {
        match[i] = -1;
        if (idx2[2 * i] >= 0 && idx2[2 * i + 1] >= 0) {
            const float d0 = (float)dist2[2 * i], d1 = (float)dist2[2 * i + 1];
            if (d0 < d1 * 0.7) match[i] = idx2[2 * i];
        }
    }
}

}  // namespace match_oracle
