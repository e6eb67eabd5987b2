// oracle/orb_oracle.cpp -- TEST INFRASTRUCTURE ONLY. See orb_oracle.h for scope/provenance.
#include "orb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <utility>

#include "cvprims.h"

namespace orb_oracle {

using cvp::cvRound;

static const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19;

const int8_t kPattern[1024] = {
#include "../orb-slam3_byzyh_b200/csrc/orb_pattern.inc"
};

// reference src/ORBextractor.cc:468-571
Extractor::Extractor(int nf, float sf, int nl, int ini, int mn)
    : nfeatures(nf), nlevels(nl), iniThFAST(ini), minThFAST(mn), scaleFactor(sf) {
    mvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvScaleFactor[0] = 1.0f;
    mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);  // float*double->float
        mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
        mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
    levels.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = (float)(nfeatures * (1 - factor) /
                             (1 - (float)pow((double)factor, (double)nlevels)));
    int sumFeatures = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        mnFeaturesPerLevel[level] = cvRound(nDesired);
        sumFeatures += mnFeaturesPerLevel[level];
        nDesired *= factor;
    }
    mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);

    umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = cvp::cvFloor(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = cvp::cvCeil(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) umax[v] = cvRound(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// reference src/ORBextractor.cc:1687-1740
void Extractor::compute_pyramid(const uint8_t* img, int rows, int cols, size_t step) {
    for (int level = 0; level < nlevels; ++level) {
        float scale = mvInvScaleFactor[level];
        Level& L = levels[level];
        L.w = cvRound((float)cols * scale);
        L.h = cvRound((float)rows * scale);
        L.step = L.w + EDGE_THRESHOLD * 2;
        L.padded.assign((size_t)L.step * (L.h + EDGE_THRESHOLD * 2), 0);
        uint8_t* roi = L.padded.data() + EDGE_THRESHOLD * L.step + EDGE_THRESHOLD;
        if (level != 0) {
            const Level& P = levels[level - 1];
            cvp::resize_linear_u8(P.roi(), P.w, P.h, P.step, roi, L.w, L.h, L.step);
            cvp::copy_make_border_reflect101(roi, L.w, L.h, L.step, L.padded.data(), L.step,
                                             EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD,
                                             EDGE_THRESHOLD);
        } else {
            cvp::copy_make_border_reflect101(img, cols, rows, step, L.padded.data(), L.step,
                                             EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD,
                                             EDGE_THRESHOLD);
        }
    }
}

// reference src/ORBextractor.cc:91-138
float ic_angle(const uint8_t* center, int step, const std::vector<int>& u_max) {
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = u_max[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return cvp::fast_atan2((float)m_01, (float)m_10);
}

// reference src/ORBextractor.cc:141, 150-203
void orb_descriptor(const uint8_t* center, int step, float angle_deg, uint8_t* desc) {
    const float factorPI = (float)(M_PI / 180.f);
    float angle = angle_deg * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            const int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2],
                      y1 = pat[4 * k + 3];
            const int t0 = center[cvRound(x0 * b + y0 * a) * step + cvRound(x0 * a - y0 * b)];
            const int t1 = center[cvRound(x1 * b + y1 * a) * step + cvRound(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ---------------------------------------------------------------------------------------
// DistributeOctTree, reference src/ORBextractor.cc:602-1057.
// ---------------------------------------------------------------------------------------
namespace {
struct Node {
    int ULx = 0, ULy = 0, URx = 0, BRy = 0;  // [ULx,URx) x [ULy,BRy)
    std::vector<int> keys;                   // candidate indices, candidate order
    std::list<Node>::iterator lit;
    bool noMore = false;
};

// :602-674
void divide(const Node& n, const std::vector<Cand>& c, Node ch[4]) {
    const int halfX = (int)ceilf((float)(n.URx - n.ULx) / 2);
    const int halfY = (int)ceilf((float)(n.BRy - n.ULy) / 2);
    const int mx = n.ULx + halfX, my = n.ULy + halfY;
    ch[0].ULx = n.ULx; ch[0].URx = mx;    ch[0].ULy = n.ULy; ch[0].BRy = my;
    ch[1].ULx = mx;    ch[1].URx = n.URx; ch[1].ULy = n.ULy; ch[1].BRy = my;
    ch[2].ULx = n.ULx; ch[2].URx = mx;    ch[2].ULy = my;    ch[2].BRy = n.BRy;
    ch[3].ULx = mx;    ch[3].URx = n.URx; ch[3].ULy = my;    ch[3].BRy = n.BRy;
    for (int k : n.keys) {
        const float px = (float)c[k].x, py = (float)c[k].y;
        if (px < mx) {
            if (py < my) ch[0].keys.push_back(k);
            else ch[2].keys.push_back(k);
        } else if (py < my) ch[1].keys.push_back(k);
        else ch[3].keys.push_back(k);
    }
    for (int i = 0; i < 4; i++)
        if (ch[i].keys.size() == 1) ch[i].noMore = true;
}

typedef std::pair<int, Node*> SizeNode;
// :676-697
bool compare_nodes(SizeNode& e1, SizeNode& e2) {
    if (e1.first < e2.first) return true;
    if (e1.first > e2.first) return false;
    return e1.second->ULx < e2.second->ULx;
}
}  // namespace

std::vector<int> distribute_octtree(const std::vector<Cand>& cands, int minX, int maxX, int minY,
                                    int maxY, int N) {
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> lNodes;
    std::vector<Node*> ini(nIni);
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * (float)i);
        ni.URx = (int)(hX * (float)(i + 1));
        ni.ULy = 0;
        ni.BRy = maxY - minY;
        lNodes.push_back(ni);
        ini[i] = &lNodes.back();
    }
    for (size_t i = 0; i < cands.size(); i++)
        ini[(int)((float)cands[i].x / hX)]->keys.push_back((int)i);

    auto lit = lNodes.begin();
    while (lit != lNodes.end()) {
        if (lit->keys.size() == 1) { lit->noMore = true; lit++; }
        else if (lit->keys.empty()) lit = lNodes.erase(lit);
        else lit++;
    }

    bool bFinish = false;
    std::vector<SizeNode> vSizeAndPointerToNode;
    auto push_children = [&](Node ch[4], int* nToExpand) {
        for (int i = 0; i < 4; i++) {
            if (ch[i].keys.size() > 0) {
                lNodes.push_front(ch[i]);
                if (ch[i].keys.size() > 1) {
                    if (nToExpand) (*nToExpand)++;
                    vSizeAndPointerToNode.push_back(
                        std::make_pair((int)ch[i].keys.size(), &lNodes.front()));
                    lNodes.front().lit = lNodes.begin();
                }
            }
        }
    };
    while (!bFinish) {
        int prevSize = (int)lNodes.size();
        lit = lNodes.begin();
        int nToExpand = 0;
        vSizeAndPointerToNode.clear();
        while (lit != lNodes.end()) {
            if (lit->noMore) { lit++; continue; }
            Node ch[4];
            divide(*lit, cands, ch);
            push_children(ch, &nToExpand);
            lit = lNodes.erase(lit);
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
            bFinish = true;
        } else if (((int)lNodes.size() + nToExpand * 3) > N) {
            while (!bFinish) {
                prevSize = (int)lNodes.size();
                std::vector<SizeNode> prev = vSizeAndPointerToNode;
                vSizeAndPointerToNode.clear();
                std::sort(prev.begin(), prev.end(), compare_nodes);
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node ch[4];
                    divide(*prev[j].second, cands, ch);
                    push_children(ch, nullptr);
                    lNodes.erase(prev[j].second->lit);
                    if ((int)lNodes.size() >= N) break;
                }
                if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
            }
        }
    }
    std::vector<int> result;
    for (auto it = lNodes.begin(); it != lNodes.end(); it++) {
        int best = it->keys[0];
        int maxResponse = cands[best].score;
        for (size_t k = 1; k < it->keys.size(); k++)
            if (cands[it->keys[k]].score > maxResponse) {
                best = it->keys[k];
                maxResponse = cands[best].score;
            }
        result.push_back(best);
    }
    return result;
}

// reference src/ORBextractor.cc:1061-1208
void Extractor::compute_keypoints_octtree() {
    const float W = 35;
    for (int level = 0; level < nlevels; ++level) {
        Level& L = levels[level];
        L.cands.clear();
        L.cell_retry.clear();
        const int minBorderX = EDGE_THRESHOLD - 3;
        const int minBorderY = minBorderX;
        const int maxBorderX = L.w - EDGE_THRESHOLD + 3;
        const int maxBorderY = L.h - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBorderX - minBorderX);
        const float height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W);
        const int nRows = (int)(height / W);
        const int wCell = (int)ceilf(width / nCols);
        const int hCell = (int)ceilf(height / nRows);
        std::vector<cvp::FastKP> cell;
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) { L.cell_retry.insert(L.cell_retry.end(), nCols, 2); continue; }
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) { L.cell_retry.push_back(2); continue; }
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const uint8_t* sub = L.roi() + (int)iniY * L.step + (int)iniX;
                const int cw = (int)maxX - (int)iniX, chh = (int)maxY - (int)iniY;
                cvp::fast9_16(sub, cw, chh, L.step, iniThFAST, true, cell);
                uint8_t retry = 0;
                if (cell.empty()) {
                    cvp::fast9_16(sub, cw, chh, L.step, minThFAST, true, cell);
                    retry = 1;
                }
                L.cell_retry.push_back(retry);
                for (const auto& k : cell)
                    L.cands.push_back({k.x + j * wCell, k.y + i * hCell, k.score});
            }
        }
        std::vector<int> keep = distribute_octtree(L.cands, minBorderX, maxBorderX, minBorderY,
                                                   maxBorderY, mnFeaturesPerLevel[level]);
        const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);
        L.kps.clear();
        for (int idx : keep) {
            OrbKp kp;
            kp.x = (float)(L.cands[idx].x + minBorderX);
            kp.y = (float)(L.cands[idx].y + minBorderY);
            kp.size = (float)scaledPatchSize;
            kp.angle = -1.f;
            kp.response = (float)L.cands[idx].score;
            kp.octave = level;
            kp.class_id = -1;
            L.kps.push_back(kp);
        }
    }
    for (int level = 0; level < nlevels; ++level) {
        Level& L = levels[level];
        for (auto& kp : L.kps)
            kp.angle = ic_angle(L.roi() + cvRound(kp.y) * L.step + cvRound(kp.x), L.step, umax);
    }
}

// reference src/ORBextractor.cc:1557-1682
int Extractor::extract(const uint8_t* img, int rows, int cols, size_t step, int lap0, int lap1,
                       std::vector<OrbKp>& out_kps, std::vector<uint8_t>& out_desc) {
    if (!img || rows <= 0 || cols <= 0) return -1;
    compute_pyramid(img, rows, cols, step);
    compute_keypoints_octtree();
    int nkeypoints = 0;
    for (int level = 0; level < nlevels; ++level) nkeypoints += (int)levels[level].kps.size();
    out_kps.assign(nkeypoints, OrbKp());
    out_desc.assign((size_t)nkeypoints * 32, 0);
    int monoIndex = 0, stereoIndex = nkeypoints - 1;
    for (int level = 0; level < nlevels; ++level) {
        Level& L = levels[level];
        if (L.kps.empty()) continue;
        L.blurred.assign((size_t)L.w * L.h, 0);
        {   // clone() of the ROI, then in-place GaussianBlur (:1629-1637)
            std::vector<uint8_t> work((size_t)L.w * L.h);
            for (int y = 0; y < L.h; y++) memcpy(&work[(size_t)y * L.w], L.roi() + y * L.step, L.w);
            cvp::gaussian_blur_7x7_s2(work.data(), L.w, L.h, L.w, L.blurred.data(), L.w);
        }
        float scale = mvScaleFactor[level];
        for (const OrbKp& k0 : L.kps) {
            OrbKp kp = k0;
            uint8_t d[32];
            orb_descriptor(L.blurred.data() + cvRound(kp.y) * L.w + cvRound(kp.x), L.w, kp.angle, d);
            if (level != 0) { kp.x *= scale; kp.y *= scale; }
            int dst;
            if (kp.x >= (float)lap0 && kp.x <= (float)lap1) dst = stereoIndex--;
            else dst = monoIndex++;
            out_kps[dst] = kp;
            memcpy(&out_desc[(size_t)dst * 32], d, 32);
        }
    }
    return monoIndex;
}

}  // namespace orb_oracle
