// search.cu -- windowed projection matching, the three Frame-based overloads of
// ORBmatcher::SearchByProjection (/root/reference/src/ORBmatcher.cc:46-240, 1951-2185, 2197-2325;
// Nleft == -1 branches) on top of Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea
// (src/Frame.cc:469-504, 962-978, 859-951) and ComputeThreeMaxima (ORBmatcher.cc:2336-2378).
//
// The reference is a sequential loop: a keypoint accepted by map point j (whose MapPoint has
// Observations() > 0) is skipped by every later map point.  Here the loop is a fixpoint of
// data-parallel passes: every pass evaluates ALL map points (one thread each, 1 M points keep
// 148 SMs busy) against "claim times" c[k] = index of the first accepted point that blocks
// keypoint k, excluding k for point j iff c[k] < j; the pass then recomputes c from its own
// acceptances.  Point j only depends on claims of points < j, so after r passes every point whose
// dependency chain is shorter than r is final and the iteration reaches exactly the sequential
// result (typically 2-3 passes: conflicts are rare and shallow).  Candidate order inside a window
// is the reference's (grid column outer, row inner, ascending keypoint index inside a cell), which
// decides ties under the strict `<`.
#include <limits.h>

#include <algorithm>
#include <vector>

#include "orbfe_internal.h"
#include "scratch.h"

namespace {

constexpr int GC = 64, GR = 48;          // FRAME_GRID_COLS / ROWS, include/Frame.h:44-45
constexpr int HISTO = 30;                // HISTO_LENGTH, ORBmatcher.cc:38
constexpr int kWarpPassMax = 1 << 16;    // up to this many map points a pass uses one warp per point

struct GridDev {
    const OrbfeKeyPoint* keys;
    const float* uright;
    const uint32_t* desc;
    int n;
    float minX, minY, maxX, maxY, wInv, hInv;
    const int* cellStart;  // [GC*GR+1], cell id = ix*GR+iy
    const int* cellItems;  // keypoint indices, ascending inside a cell
};

struct PtsDev {
    int m;
    int jOff;   // global index of point 0: claims and assignments speak global indices (a map shard starts above 0)
    const float *u, *v, *ur, *radius, *angle;
    const int *minLevel, *maxLevel;
    const uint8_t *valid, *blocks;
    const uint32_t* desc;
};

// AssignFeaturesToGrid: one CTA; count -> scan -> fill -> per-cell insertion sort (ascending index
// == the reference's push_back order).
__global__ void __launch_bounds__(1024)
k_build_grid(const OrbfeKeyPoint* __restrict__ keys, int n, float minX, float minY, float wInv, float hInv,
             int* __restrict__ cellOf, int* __restrict__ cellStart, int* __restrict__ cellItems) {
    __shared__ int cnt[GC * GR];
    __shared__ int wsum[32];
    const int tid = threadIdx.x;
    for (int c = tid; c < GC * GR; c += 1024) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int px = (int)roundf((keys[i].x - minX) * wInv);  // PosInGrid, Frame.cc:967-968
        const int py = (int)roundf((keys[i].y - minY) * hInv);
        int c = -1;
        if (px >= 0 && px < GC && py >= 0 && py < GR) {
            c = px * GR + py;
            atomicAdd(&cnt[c], 1);
        }
        cellOf[i] = c;
    }
    __syncthreads();
    // exclusive scan of 3072 counts: 3 per thread
    int v[3], s = 0;
    for (int k = 0; k < 3; k++) { v[k] = cnt[tid * 3 + k]; s += v[k]; }
    int incl = s;
    const int lane = tid & 31, wid = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    int base = incl - s + (wid ? wsum[wid - 1] : 0);
    for (int k = 0; k < 3; k++) { cellStart[tid * 3 + k] = base; cnt[tid * 3 + k] = base; base += v[k]; }
    if (tid == 1023) cellStart[GC * GR] = base;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int c = cellOf[i];
        if (c >= 0) cellItems[atomicAdd(&cnt[c], 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < GC * GR; c += 1024) {
        const int b = cellStart[c], e = cnt[c];
        for (int i = b + 1; i < e; i++) {
            const int x = cellItems[i];
            int j = i - 1;
            while (j >= b && cellItems[j] > x) { cellItems[j + 1] = cellItems[j]; j--; }
            cellItems[j + 1] = x;
        }
    }
}

__device__ __forceinline__ int hamming8(const uint32_t* a, const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// One pass over all map points.  claimIn[k] < j  <=>  keypoint k is taken for point j.
__global__ void __launch_bounds__(128)
k_search_pass(GridDev F, PtsDev P, int mode, int thAccept, float nnratio, const int* __restrict__ claimIn,
              int* __restrict__ claimOut, int* __restrict__ bestIdx, int* __restrict__ bestDist) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.m) return;
    int bD = 256, bL = -1, bD2 = 256, bL2 = -1, bI = -1;
    if (P.valid[j]) {
        const float x = P.u[j], y = P.v[j], r = P.radius[j];
        const int minLevel = P.minLevel[j], maxLevel = P.maxLevel[j];
        // GetFeaturesInArea, Frame.cc:869-906
        const int c0x = max(0, (int)floorf((x - F.minX - r) * F.wInv));
        const int c1x = min(GC - 1, (int)ceilf((x - F.minX + r) * F.wInv));
        const int c0y = max(0, (int)floorf((y - F.minY - r) * F.hInv));
        const int c1y = min(GR - 1, (int)ceilf((y - F.minY + r) * F.hInv));
        if (c0x < GC && c1x >= 0 && c0y < GR && c1y >= 0) {
            const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
            uint32_t d[8];
            const uint4* pd = reinterpret_cast<const uint4*>(P.desc + 8 * (size_t)j);
            *reinterpret_cast<uint4*>(d) = pd[0];
            *reinterpret_cast<uint4*>(d + 4) = pd[1];
            const float ur = P.ur ? P.ur[j] : 0.f;
            for (int ix = c0x; ix <= c1x; ix++)
                for (int iy = c0y; iy <= c1y; iy++) {
                    const int cb = F.cellStart[ix * GR + iy], ce = F.cellStart[ix * GR + iy + 1];
                    for (int t = cb; t < ce; t++) {
                        const int idx = F.cellItems[t];
                        const OrbfeKeyPoint kp = F.keys[idx];
                        if (checkLevels) {
                            if (kp.octave < minLevel) continue;
                            if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                        }
                        if (!(fabsf(kp.x - x) < r && fabsf(kp.y - y) < r)) continue;
                        if (claimIn[idx] < j + P.jOff) continue;  // ORBmatcher.cc:103-105 / 2040-2042 / 2266-2267
                        if (mode != ORBFE_SEARCH_KEYFRAME && F.uright) {
                            const float uR = F.uright[idx];
                            if (uR > 0 && fabsf(ur - uR) > r) continue;  // :108-118 / :2044-2050
                        }
                        const uint4* kd = reinterpret_cast<const uint4*>(F.desc + 8 * (size_t)idx);
                        const int dist = hamming8(d, kd[0], kd[1]);
                        if (dist < bD) {
                            bD2 = bD; bD = dist; bL2 = bL; bL = kp.octave; bI = idx;
                        } else if (mode == ORBFE_SEARCH_MAPPOINTS && dist < bD2) {
                            bL2 = kp.octave; bD2 = dist;
                        }
                    }
                }
        }
    }
    bool accept = bI >= 0 && bD <= thAccept;
    if (accept && mode == ORBFE_SEARCH_MAPPOINTS && bL == bL2 && (float)bD > nnratio * (float)bD2) accept = false;
    bestDist[j] = bD;
    bestIdx[j] = accept ? bI : -1;
    if (accept && (mode == ORBFE_SEARCH_KEYFRAME || (P.blocks ? P.blocks[j] != 0 : true))) atomicMin(&claimOut[bI], j + P.jOff);
}

// The same pass with one WARP per map point, for the per-frame calls of Tracking (a few thousand points: one thread
// per point leaves the machine empty and walks ~50 candidates serially).  Lanes stride over the candidates of each
// grid column; best and second best are the two smallest keys (distance << 23 | position in GetFeaturesInArea order),
// which is exactly what the sequential update rule (:127-149) leaves behind: a later candidate of equal distance
// becomes the second best, a new best demotes the previous one.
__global__ void __launch_bounds__(128)
k_search_pass_warp(GridDev F, PtsDev P, int mode, int thAccept, float nnratio, const int* __restrict__ claimIn,
                   int* __restrict__ claimOut, int* __restrict__ bestIdx, int* __restrict__ bestDist) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= P.m) return;
    uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    int i1 = -1, i2 = -1;
    if (P.valid[j]) {
        const float x = P.u[j], y = P.v[j], r = P.radius[j];
        const int minLevel = P.minLevel[j], maxLevel = P.maxLevel[j];
        const int c0x = max(0, (int)floorf((x - F.minX - r) * F.wInv));
        const int c1x = min(GC - 1, (int)ceilf((x - F.minX + r) * F.wInv));
        const int c0y = max(0, (int)floorf((y - F.minY - r) * F.hInv));
        const int c1y = min(GR - 1, (int)ceilf((y - F.minY + r) * F.hInv));
        if (c0x < GC && c1x >= 0 && c0y < GR && c1y >= 0) {
            const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
            uint32_t d[8];
            const uint4* pd = reinterpret_cast<const uint4*>(P.desc + 8 * (size_t)j);
            *reinterpret_cast<uint4*>(d) = pd[0];
            *reinterpret_cast<uint4*>(d + 4) = pd[1];
            const float ur = P.ur ? P.ur[j] : 0.f;
            int order = 0;
            for (int ix = c0x; ix <= c1x; ix++) {
                const int cb = F.cellStart[ix * GR + c0y], ce = F.cellStart[ix * GR + c1y + 1];
                for (int t = cb + lane; t < ce; t += 32) {
                    const int idx = F.cellItems[t];
                    const OrbfeKeyPoint kp = F.keys[idx];
                    if (checkLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    if (!(fabsf(kp.x - x) < r && fabsf(kp.y - y) < r)) continue;
                    if (claimIn[idx] < j + P.jOff) continue;
                    if (mode != ORBFE_SEARCH_KEYFRAME && F.uright) {
                        const float uR = F.uright[idx];
                        if (uR > 0 && fabsf(ur - uR) > r) continue;
                    }
                    const uint4* kd = reinterpret_cast<const uint4*>(F.desc + 8 * (size_t)idx);
                    const uint32_t k = ((uint32_t)hamming8(d, kd[0], kd[1]) << 23) | (uint32_t)(order + t - cb);
                    if (k < k1) { k2 = k1; i2 = i1; k1 = k; i1 = idx; }
                    else if (k < k2) { k2 = k; i2 = idx; }
                }
                order += ce - cb;
            }
        }
    }
    // warp top-2 of unique keys, carrying the keypoint index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        const int j1 = __shfl_xor_sync(0xffffffffu, i1, o), j2 = __shfl_xor_sync(0xffffffffu, i2, o);
        uint32_t hi; int hiI;
        if (o1 < k1) { hi = k1; hiI = i1; k1 = o1; i1 = j1; } else { hi = o1; hiI = j1; }
        // second = min(hi, k2, o2)
        if (o2 < k2) { k2 = o2; i2 = j2; }
        if (hi < k2) { k2 = hi; i2 = hiI; }
    }
    if (lane != 0) return;
    const int bD = (k1 == 0xFFFFFFFFu || (k1 >> 23) >= 256) ? 256 : (int)(k1 >> 23);
    const int bI = bD < 256 ? i1 : -1;
    int bD2 = 256, bL2 = -1;
    if (mode == ORBFE_SEARCH_MAPPOINTS && k2 != 0xFFFFFFFFu && (k2 >> 23) < 256) { bD2 = (int)(k2 >> 23); bL2 = F.keys[i2].octave; }
    const int bL = bI >= 0 ? F.keys[bI].octave : -1;
    bool accept = bI >= 0 && bD <= thAccept;
    if (accept && mode == ORBFE_SEARCH_MAPPOINTS && bL == bL2 && (float)bD > nnratio * (float)bD2) accept = false;
    bestDist[j] = bD;
    bestIdx[j] = accept ? bI : -1;
    if (accept && (mode == ORBFE_SEARCH_KEYFRAME || (P.blocks ? P.blocks[j] != 0 : true))) atomicMin(&claimOut[bI], j + P.jOff);
}

__global__ void k_claims_init(const uint8_t* __restrict__ claimed, int n, int* __restrict__ a, int* __restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = (claimed && claimed[i]) ? -1 : INT_MAX;
    a[i] = v;
    b[i] = v;
}

// changed |= (a != b); then b (the older table) is reset to the static claims for the next pass.
__global__ void k_claims_diff(const int* __restrict__ a, int* __restrict__ b, const uint8_t* __restrict__ claimed,
                              int n, int* __restrict__ changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (a[i] != b[i]) *changed = 1;
    b[i] = (claimed && claimed[i]) ? -1 : INT_MAX;
}

// Final bookkeeping: assignments (last accepted point wins), rotation histogram votes.
__global__ void k_search_assign(GridDev F, PtsDev P, int useHist, const int* __restrict__ bestIdx,
                                int* __restrict__ assigned, int* __restrict__ hist, int* __restrict__ binOf,
                                int* __restrict__ nmatches) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.m) return;
    const int k = bestIdx[j];
    if (k < 0) return;
    atomicMax(&assigned[k], j + P.jOff);
    atomicAdd(nmatches, 1);
    if (useHist) {
        float rot = P.angle[j] - F.keys[k].angle;  // ORBmatcher.cc:2074-2084
        if (rot < 0.0f) rot += 360.0f;
        int bin = (int)roundf(rot * (1.0f / HISTO));
        if (bin == HISTO) bin = 0;
        bin = min(max(bin, 0), HISTO - 1);
        binOf[j] = bin;
        atomicAdd(&hist[bin], 1);
    }
}

__global__ void k_search_cull(int m, const int* __restrict__ bestIdx, const int* __restrict__ binOf,
                              const int* __restrict__ hist, int* __restrict__ assigned, int* __restrict__ nmatches) {
    __shared__ int keep[3];
    if (threadIdx.x == 0) {  // ComputeThreeMaxima, ORBmatcher.cc:2336-2378
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int k = bestIdx[j];
    if (k < 0) return;
    const int b = binOf[j];
    if (b != keep[0] && b != keep[1] && b != keep[2]) {
        assigned[k] = -1;        // every write stores the same value: no race on the result
        atomicSub(nmatches, 1);
    }
}

// atomicMax above must not lose against a concurrent -1 store of the cull pass: the cull runs in
// its own launch after k_search_assign.  Keypoints never touched keep their input value because
// k_search_assign only raises entries it owns; entries that held a larger input index are fixed up:
__global__ void k_assign_prepare(const int* __restrict__ bestIdx, int m, int* __restrict__ assigned) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int k = bestIdx[j];
    if (k >= 0) assigned[k] = -1;  // will be overwritten by atomicMax(j >= 0)
}

// ---------------------------------------------------------------------------------------------
// Fisheye stereo frames (F.Nleft != -1): every map point is searched in the left camera and then
// in the right camera (ORBmatcher.cc:171-237, 2090-2155); F.mvpMapPoints slots are [0,Nleft) for
// the left and [Nleft,N) for the right keypoints, and in the map-point overload an accepted match
// also writes its stereo partner's slot (mvLeftToRightMatch / mvRightToLeftMatch, :159-163,
// :215-219).  Claim times are 2j (left search of point j) and 2j+1 (right search), so the right
// search of a point sees what its own left search claimed.
struct FeDev {
    GridDev L, R;
    const int* l2r;
    const int* r2l;
    int Nl;
};
struct PtsFe {
    PtsDev L;                      // left projections + shared angle / blocks / desc
    const float *u, *v, *radius;   // right projections
    const int *minLevel, *maxLevel;
    const uint8_t* valid;
};

struct WinBest { int bD, bL, bD2, bL2, bI; bool had; };

__device__ __forceinline__ WinBest win_search(const GridDev& F, int slotOff, float x, float y, float r, int minLevel,
                                              int maxLevel, const uint32_t* d, const int* __restrict__ claimIn, int t,
                                              bool second) {
    WinBest w = {256, -1, 256, -1, -1, false};
    const int c0x = max(0, (int)floorf((x - F.minX - r) * F.wInv));
    const int c1x = min(GC - 1, (int)ceilf((x - F.minX + r) * F.wInv));
    const int c0y = max(0, (int)floorf((y - F.minY - r) * F.hInv));
    const int c1y = min(GR - 1, (int)ceilf((y - F.minY + r) * F.hInv));
    if (!(c0x < GC && c1x >= 0 && c0y < GR && c1y >= 0)) return w;
    const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = c0x; ix <= c1x; ix++)
        for (int iy = c0y; iy <= c1y; iy++) {
            const int cb = F.cellStart[ix * GR + iy], ce = F.cellStart[ix * GR + iy + 1];
            for (int q = cb; q < ce; q++) {
                const int idx = F.cellItems[q];
                const OrbfeKeyPoint kp = F.keys[idx];
                if (checkLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                if (!(fabsf(kp.x - x) < r && fabsf(kp.y - y) < r)) continue;
                w.had = true;                                    // vIndices is not empty
                if (claimIn[idx + slotOff] < t) continue;
                const uint4* kd = reinterpret_cast<const uint4*>(F.desc + 8 * (size_t)idx);
                const int dist = hamming8(d, kd[0], kd[1]);
                if (dist < w.bD) { w.bD2 = w.bD; w.bD = dist; w.bL2 = w.bL; w.bL = kp.octave; w.bI = idx; }
                else if (second && dist < w.bD2) { w.bL2 = kp.octave; w.bD2 = dist; }
            }
        }
    return w;
}

__global__ void __launch_bounds__(128)
k_search_pass_fe(FeDev F, PtsFe P, int mode, int thAccept, float nnratio, const int* __restrict__ claimIn,
                 int* __restrict__ claimOut, int* __restrict__ bestL, int* __restrict__ bestR) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.L.m) return;
    uint32_t d[8];
    const uint4* pd = reinterpret_cast<const uint4*>(P.L.desc + 8 * (size_t)j);
    *reinterpret_cast<uint4*>(d) = pd[0];
    *reinterpret_cast<uint4*>(d + 4) = pd[1];
    const bool blocks = P.L.blocks ? P.L.blocks[j] != 0 : true;
    const bool m0 = mode == ORBFE_SEARCH_MAPPOINTS;
    int bl = -1, br = -1;
    bool doRight = m0 ? P.valid[j] != 0 : true;   // :2090 the last-frame overload always looks right
    if (P.L.valid[j]) {
        const WinBest w = win_search(F.L, 0, P.L.u[j], P.L.v[j], P.L.radius[j], P.L.minLevel[j], P.L.maxLevel[j], d,
                                     claimIn, 2 * j, m0);
        if (w.had && w.bD <= thAccept) {
            if (m0 && w.bL == w.bL2 && (float)w.bD > nnratio * (float)w.bD2) doRight = false;   // :154 `continue`
            else bl = w.bI;
        }
        if (!m0 && !w.had) doRight = false;                                                      // :2027 `continue`
    } else if (!m0) {
        doRight = false;
    }
    if (bl >= 0 && blocks) {
        atomicMin(&claimOut[bl], 2 * j);
        if (m0 && F.l2r[bl] != -1) atomicMin(&claimOut[F.Nl + F.l2r[bl]], 2 * j);
    }
    if (doRight) {
        // this point's own left-camera claims carry time 2j < 2j+1: at the fixpoint they are in claimIn
        const WinBest w = win_search(F.R, F.Nl, P.u[j], P.v[j], P.radius[j], P.minLevel[j], P.maxLevel[j], d, claimIn,
                                     2 * j + 1, m0);
        if (w.had && w.bD <= thAccept && !(m0 && w.bL == w.bL2 && (float)w.bD > nnratio * (float)w.bD2)) br = w.bI;
        if (br >= 0 && blocks) {
            atomicMin(&claimOut[F.Nl + br], 2 * j + 1);
            if (m0 && F.r2l[br] != -1) atomicMin(&claimOut[F.r2l[br]], 2 * j + 1);
        }
    }
    bestL[j] = bl;
    bestR[j] = br;
}

// Slot writes in time order: the latest write wins (tmax holds the time of the latest write).
__global__ void k_fe_assign(FeDev F, PtsFe P, int mode, int useHist, const int* __restrict__ bestL,
                            const int* __restrict__ bestR, int* __restrict__ tmax, int* __restrict__ hist,
                            int* __restrict__ binL, int* __restrict__ binR, int* __restrict__ nmatches) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.L.m) return;
    const bool m0 = mode == ORBFE_SEARCH_MAPPOINTS;
    int n = 0;
    const int bl = bestL[j], br = bestR[j];
    const float factor = 1.0f / HISTO;
    if (bl >= 0) {
        atomicMax(&tmax[bl], 2 * j); n++;
        if (m0 && F.l2r[bl] != -1) { atomicMax(&tmax[F.Nl + F.l2r[bl]], 2 * j); n++; }
        if (useHist) {
            float rot = P.L.angle[j] - F.L.keys[bl].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            bin = min(max(bin, 0), HISTO - 1);
            binL[j] = bin;
            atomicAdd(&hist[bin], 1);
        }
    }
    if (br >= 0) {
        if (m0 && F.r2l[br] != -1) { atomicMax(&tmax[F.r2l[br]], 2 * j + 1); n++; }
        atomicMax(&tmax[F.Nl + br], 2 * j + 1); n++;
        if (useHist) {
            float rot = P.L.angle[j] - F.R.keys[br].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            bin = min(max(bin, 0), HISTO - 1);
            binR[j] = bin;
            atomicAdd(&hist[bin], 1);
        }
    }
    if (n) atomicAdd(nmatches, n);
}

__global__ void k_fe_commit(int N, const int* __restrict__ tmax, int* __restrict__ assigned) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N && tmax[k] >= 0) assigned[k] = tmax[k] >> 1;
}

__global__ void k_fe_cull(int m, int Nl, const int* __restrict__ bestL, const int* __restrict__ bestR,
                          const int* __restrict__ binL, const int* __restrict__ binR, const int* __restrict__ hist,
                          int* __restrict__ assigned, int* __restrict__ nmatches) {
    __shared__ int keep[3];
    if (threadIdx.x == 0) {  // ComputeThreeMaxima, ORBmatcher.cc:2336-2378
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    int n = 0;
    if (bestL[j] >= 0) {
        const int b = binL[j];
        if (b != keep[0] && b != keep[1] && b != keep[2]) { assigned[bestL[j]] = -1; n++; }
    }
    if (bestR[j] >= 0) {
        const int b = binR[j];
        if (b != keep[0] && b != keep[1] && b != keep[2]) { assigned[Nl + bestR[j]] = -1; n++; }
    }
    if (n) atomicSub(nmatches, n);
}

// Exact sequential form (one thread): used when a non-blocking point could overwrite -- and thereby
// un-block -- a slot through a stereo partner write, which the claim-time fixpoint cannot express.
__global__ void k_search_fe_sequential(FeDev F, PtsFe P, int mode, int thAccept, float nnratio,
                                       int* __restrict__ claim /* INT_MAX = free, -1 = taken */,
                                       int* __restrict__ assigned, int* __restrict__ bestL, int* __restrict__ bestR,
                                       int* __restrict__ nmatches) {
    if (blockIdx.x | threadIdx.x) return;
    const bool m0 = mode == ORBFE_SEARCH_MAPPOINTS;
    int n = 0;
    for (int j = 0; j < P.L.m; j++) {
        uint32_t d[8];
        for (int i = 0; i < 8; i++) d[i] = P.L.desc[8 * (size_t)j + i];
        const int taken = P.L.blocks ? (P.L.blocks[j] ? -1 : INT_MAX) : -1;
        int bl = -1, br = -1;
        bool doRight = m0 ? P.valid[j] != 0 : true;   // :2090 the last-frame overload always looks right
        if (P.L.valid[j]) {
            const WinBest w = win_search(F.L, 0, P.L.u[j], P.L.v[j], P.L.radius[j], P.L.minLevel[j], P.L.maxLevel[j],
                                         d, claim, 0, m0);
            if (w.had && w.bD <= thAccept) {
                if (m0 && w.bL == w.bL2 && (float)w.bD > nnratio * (float)w.bD2) doRight = false;
                else bl = w.bI;
            }
            if (!m0 && !w.had) doRight = false;
        } else if (!m0) {
            doRight = false;
        }
        if (bl >= 0) {
            assigned[bl] = j; claim[bl] = taken; n++;
            if (m0 && F.l2r[bl] != -1) { assigned[F.Nl + F.l2r[bl]] = j; claim[F.Nl + F.l2r[bl]] = taken; n++; }
        }
        if (doRight) {
            const WinBest w = win_search(F.R, F.Nl, P.u[j], P.v[j], P.radius[j], P.minLevel[j], P.maxLevel[j], d,
                                         claim, 0, m0);
            if (w.had && w.bD <= thAccept && !(m0 && w.bL == w.bL2 && (float)w.bD > nnratio * (float)w.bD2)) br = w.bI;
            if (br >= 0) {
                if (m0 && F.r2l[br] != -1) { assigned[F.r2l[br]] = j; claim[F.r2l[br]] = taken; n++; }
                assigned[F.Nl + br] = j; claim[F.Nl + br] = taken; n++;
            }
        }
        bestL[j] = bl;
        bestR[j] = br;
    }
    *nmatches = n;
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::SearchForInitialization (ORBmatcher.cc:735-891).  A keypoint of F2 can be taken over by a
// later keypoint of F1 with a strictly smaller distance (vMatchedDistance), so the loop over F1 is
// kept in order; ONE warp walks it and spreads the ~100 window candidates of each step over its
// lanes.  Candidate order (ties under the strict `<`) is carried in the low bits of the packed key.
__global__ void __launch_bounds__(32)
k_search_init(const OrbfeKeyPoint* __restrict__ keys1, const uint32_t* __restrict__ desc1, int n1, GridDev F2,
              float* __restrict__ prev, float window, float nnratio, int checkOri, int thLow,
              int* __restrict__ m12, int* __restrict__ m21, int* __restrict__ matchedDist, int* __restrict__ binOf,
              int* __restrict__ nmatchesOut) {
    const int lane = threadIdx.x;
    __shared__ int hist[HISTO];
    for (int i = lane; i < HISTO; i += 32) hist[i] = 0;
    for (int i = lane; i < n1; i += 32) { m12[i] = -1; binOf[i] = -1; }
    for (int i = lane; i < F2.n; i += 32) { m21[i] = -1; matchedDist[i] = INT_MAX; }
    __syncwarp();
    int nmatches = 0;
    const float factor = 1.0f / HISTO;
    for (int i1 = 0; i1 < n1; i1++) {
        const OrbfeKeyPoint kp1 = keys1[i1];
        if (kp1.octave > 0) continue;
        const float x = prev[2 * i1], y = prev[2 * i1 + 1], r = window;
        const int c0x = max(0, (int)floorf((x - F2.minX - r) * F2.wInv));
        const int c1x = min(GC - 1, (int)ceilf((x - F2.minX + r) * F2.wInv));
        const int c0y = max(0, (int)floorf((y - F2.minY - r) * F2.hInv));
        const int c1y = min(GR - 1, (int)ceilf((y - F2.minY + r) * F2.hInv));
        if (!(c0x < GC && c1x >= 0 && c0y < GR && c1y >= 0)) continue;
        uint32_t d[8];
        const uint4* pd = reinterpret_cast<const uint4*>(desc1 + 8 * (size_t)i1);
        *reinterpret_cast<uint4*>(d) = pd[0];
        *reinterpret_cast<uint4*>(d + 4) = pd[1];
        uint32_t b0 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu;   // packed (dist << 20 | order), order < 2^20
        bool had = false;
        int seq = 0;
        for (int ix = c0x; ix <= c1x; ix++) {
            // the cells (ix, c0y..c1y) are consecutive in cellItems: one contiguous run per grid column
            const int cb = F2.cellStart[ix * GR + c0y], ce = F2.cellStart[ix * GR + c1y + 1];
            for (int tb = cb; tb < ce; tb += 32, seq += 32) {
                const int t = tb + lane;
                if (t < ce) {
                    const int i2 = F2.cellItems[t];
                    const OrbfeKeyPoint kp = F2.keys[i2];
                    // level1 == 0: GetFeaturesInArea(.., 0, 0) keeps octave 0 only (bCheckLevels is true)
                    if (kp.octave == 0 && fabsf(kp.x - x) < r && fabsf(kp.y - y) < r) {
                        had = true;
                        const uint4* kd = reinterpret_cast<const uint4*>(F2.desc + 8 * (size_t)i2);
                        const int dist = hamming8(d, kd[0], kd[1]);
                        if (!(matchedDist[i2] <= dist)) {
                            const uint32_t key = ((uint32_t)dist << 20) | (uint32_t)(seq + lane);
                            const uint32_t hi = max(b0, key);
                            b0 = min(b0, key);
                            b1 = min(b1, hi);
                        }
                    }
                }
            }
        }
        had = __any_sync(0xffffffffu, had);
        if (!had) continue;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t o0 = __shfl_xor_sync(0xffffffffu, b0, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o);
            uint32_t hi = max(b0, o0);
            b0 = min(b0, o0);
            b1 = min(b1, hi);
            hi = max(b0, o1);   // o1 >= o0 >= new b0 is not guaranteed after the first min: insert again
            b0 = min(b0, o1);
            b1 = min(b1, hi);
        }
        if (b0 == 0xFFFFFFFFu) continue;                       // bestDist == INT_MAX
        const int bestDist = (int)(b0 >> 20);
        const float second = b1 == 0xFFFFFFFFu ? (float)INT_MAX : (float)(int)(b1 >> 20);
        if (bestDist <= thLow && (float)bestDist < second * nnratio) {
            // which keypoint carries order (b0 & 0xFFFFF)?  recompute its index from the run layout
            int best = -1;
            {
                int want = (int)(b0 & 0xFFFFFu), s = 0;
                for (int ix = c0x; ix <= c1x && best < 0; ix++) {
                    const int cb = F2.cellStart[ix * GR + c0y], ce = F2.cellStart[ix * GR + c1y + 1];
                    const int len = ce - cb, padded = (len + 31) & ~31;
                    if (want < s + padded) best = F2.cellItems[cb + (want - s)];
                    s += padded;
                }
            }
            if (lane == 0) {
                if (m21[best] >= 0) { m12[m21[best]] = -1; nmatches--; }
                m12[i1] = best;
                m21[best] = i1;
                matchedDist[best] = bestDist;
                nmatches++;
                if (checkOri) {
                    float rot = kp1.angle - F2.keys[best].angle;
                    if (rot < 0.0f) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO) bin = 0;
                    bin = min(max(bin, 0), HISTO - 1);
                    binOf[i1] = bin;
                    hist[bin]++;
                }
            }
            __syncwarp();   // the next steps read matchedDist / m21 written by lane 0
        }
    }
    __syncwarp();
    if (checkOri) {
        int keep0, keep1, keep2;
        {
            int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
            for (int i = 0; i < HISTO; i++) {
                const int s = hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
            else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
            keep0 = ind1; keep1 = ind2; keep2 = ind3;
        }
        int removed = 0;
        for (int i = lane; i < n1; i += 32) {
            const int b = binOf[i];
            if (b >= 0 && b != keep0 && b != keep1 && b != keep2 && m12[i] >= 0) { m12[i] = -1; removed++; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;     // lane 0 holds the running count; every lane now knows `removed`
    }
    __syncwarp();
    for (int i = lane; i < n1; i += 32)
        if (m12[i] >= 0) { prev[2 * i] = F2.keys[m12[i]].x; prev[2 * i + 1] = F2.keys[m12[i]].y; }
    if (lane == 0) *nmatchesOut = nmatches;
}

// ---------------------------------------------------------------------------------------------
// Keyframe-side searches: ORBmatcher::Fuse (ORBmatcher.cc:1326-1534, 1536-1688),
// SearchBySim3 (:1690-1940) and the inner loop of the Sim3 SearchByProjection overloads (:496-733)
// carry no state from one map point to the next: KeyFrame::GetFeaturesInArea (KeyFrame.cc:843-892,
// the same cells and order as Frame's), the `kpLevel < nPredictedLevel-1 || kpLevel > nPredictedLevel`
// filter, optionally Fuse's reprojection gate (:1436-1461), then the strict-`<` Hamming argmin.
// One WARP per point (these calls carry a few thousand points: latency, not throughput): lanes
// stride over the candidates of each cell; the key (distance << 23 | position in GetFeaturesInArea
// order) makes the warp minimum the candidate the sequential loop would have kept.
__global__ void __launch_bounds__(128)
k_window_best(GridDev F, PtsDev P, int thAccept, int fuseGate, const float* __restrict__ invSigma2, int nLevels,
              int* __restrict__ bestIdx, int* __restrict__ bestDist) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= P.m) return;
    uint32_t key = 0xFFFFFFFFu;
    int kIdx = -1;
    if (P.valid[j]) {
        const float x = P.u[j], y = P.v[j], r = P.radius[j];
        const int minLevel = P.minLevel[j], maxLevel = P.maxLevel[j];
        const int c0x = max(0, (int)floorf((x - F.minX - r) * F.wInv));
        const int c1x = min(GC - 1, (int)ceilf((x - F.minX + r) * F.wInv));
        const int c0y = max(0, (int)floorf((y - F.minY - r) * F.hInv));
        const int c1y = min(GR - 1, (int)ceilf((y - F.minY + r) * F.hInv));
        if (c0x < GC && c1x >= 0 && c0y < GR && c1y >= 0) {
            uint32_t d[8];
            const uint4* pd = reinterpret_cast<const uint4*>(P.desc + 8 * (size_t)j);
            *reinterpret_cast<uint4*>(d) = pd[0];
            *reinterpret_cast<uint4*>(d + 4) = pd[1];
            const float ur = P.ur ? P.ur[j] : 0.f;
            int order = 0;
            for (int ix = c0x; ix <= c1x; ix++) {
                // the cells (ix, c0y..c1y) are contiguous in cellStart (cell id = ix*GR + iy)
                const int cb = F.cellStart[ix * GR + c0y], ce = F.cellStart[ix * GR + c1y + 1];
                for (int t = cb + lane; t < ce; t += 32) {
                    const int idx = F.cellItems[t];
                    const OrbfeKeyPoint kp = F.keys[idx];
                    if (!(fabsf(kp.x - x) < r && fabsf(kp.y - y) < r)) continue;
                    if (kp.octave < minLevel || kp.octave > maxLevel) continue;
                    if (fuseGate) {
                        if (kp.octave < 0 || kp.octave >= nLevels) continue;
                        const float ex = x - kp.x, ey = y - kp.y;
                        const float inv = invSigma2[kp.octave];
                        const float uR = F.uright ? F.uright[idx] : -1.f;
                        if (uR >= 0) {
                            const float er = ur - uR;
                            const float e2 = ex * ex + ey * ey + er * er;
                            if ((double)(e2 * inv) > 7.8) continue;
                        } else {
                            const float e2 = ex * ex + ey * ey;
                            if ((double)(e2 * inv) > 5.99) continue;
                        }
                    }
                    const uint4* kd = reinterpret_cast<const uint4*>(F.desc + 8 * (size_t)idx);
                    const int dist = hamming8(d, kd[0], kd[1]);
                    const uint32_t k = ((uint32_t)dist << 23) | (uint32_t)(order + t - cb);
                    if (k < key) { key = k; kIdx = idx; }
                }
                order += ce - cb;
            }
        }
    }
    const uint32_t best = __reduce_min_sync(0xffffffffu, key);
    const uint32_t owner = __ballot_sync(0xffffffffu, key == best && kIdx >= 0);
    if (owner == 0) {   // no candidate, or only candidates at distance 256 (never accepted by any caller)
        if (lane == 0) { bestIdx[j] = -1; bestDist[j] = 256; }
        return;
    }
    if (lane == __ffs(owner) - 1) {
        const int dist = (int)(best >> 23);
        bestDist[j] = dist < 256 ? dist : 256;
        bestIdx[j] = (dist < 256 && dist <= thAccept) ? kIdx : -1;
    }
}

// SearchBySim3, ORBmatcher.cc:1922-1937: keep i1 -> idx2 only when KF2's point at idx2 chose i1.
__global__ void k_sim3_mutual(int n1, const int* __restrict__ m1, const int* __restrict__ m2, int* __restrict__ match12,
                              int* __restrict__ nFound) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i1 >= n1) return;
    const int idx2 = m1[i1];
    int out = -1;
    if (idx2 >= 0 && m2[idx2] == i1) { out = idx2; atomicAdd(nFound, 1); }
    match12[i1] = out;
}

int sfail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define SCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return sfail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

}  // namespace

extern "C" int orbfe_search_by_projection(const OrbfeFrameView* frame, const OrbfeProjPoints* pts,
                                          const OrbfeSearchParams* prm, const uint8_t* claimed,
                                          int32_t* assigned, int32_t* best_idx, int32_t* best_dist, int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return sfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return sfail(ORBFE_ERR_INVALID, "bad device ordinal");
    if (!frame || !pts || !prm || !assigned) return sfail(ORBFE_ERR_INVALID, "null argument");
    const int n = frame->n, m = pts->m;
    if (n < 0 || m < 0 || prm->mode < 0 || prm->mode > 2) return sfail(ORBFE_ERR_INVALID, "bad sizes or mode");
    if (m == 0 || n == 0) {
        for (int j = 0; j < m; j++) {
            if (best_idx) best_idx[j] = -1;
            if (best_dist) best_dist[j] = 256;
        }
        return 0;
    }
    const bool useHist = prm->mode != ORBFE_SEARCH_MAPPOINTS && prm->check_orientation;
    if (!frame->keys || !frame->desc || !pts->u || !pts->v || !pts->radius || !pts->min_level || !pts->max_level ||
        !pts->desc || (useHist && !pts->angle))
        return sfail(ORBFE_ERR_INVALID, "missing frame / map-point array");

    // one slab: inputs | assigned (in/out) | work | outputs   (scratch.h)
    OrbfeStage S;
    std::vector<uint8_t> ones;
    if (!pts->valid) ones.assign(m, 1);
    const size_t N4 = 4 * (size_t)n, M4 = 4 * (size_t)m;
    const size_t iKeys = S.in(frame->keys, sizeof(OrbfeKeyPoint) * (size_t)n), iDesc = S.in(frame->desc, 32 * (size_t)n);
    const size_t iUr = S.in(frame->uright, frame->uright ? N4 : 0);
    const size_t iU = S.in(pts->u, M4), iV = S.in(pts->v, M4), iPur = S.in(pts->ur, pts->ur ? M4 : 0);
    const size_t iRad = S.in(pts->radius, M4), iAng = S.in(pts->angle, pts->angle ? M4 : 0);
    const size_t iMin = S.in(pts->min_level, M4), iMax = S.in(pts->max_level, M4);
    const size_t iVal = S.in(pts->valid ? pts->valid : ones.data(), (size_t)m), iBlk = S.in(pts->blocks, pts->blocks ? (size_t)m : 0);
    const size_t iPd = S.in(pts->desc, 32 * (size_t)m), iCl = S.in(claimed, claimed ? (size_t)n : 0);
    const size_t ioAsg = S.inout(assigned, N4);
    const size_t wCellOf = S.work(N4), wStart = S.work(4 * (GC * GR + 1)), wItems = S.work(N4), wA = S.work(N4), wB = S.work(N4);
    const size_t wBin = S.work(M4), wFlag = S.work(256);
    int nmatches = 0;
    const size_t oBest = S.out(best_idx, M4), oDist = S.out(best_dist, M4), oHist = S.work(4 * (HISTO + 2)), oN = S.out(&nmatches, 4);
    SCK(S.commit(device));
    cudaStream_t st = S.stream();
    SCK(S.upload());

    GridDev F;
    F.keys = S.ptr<OrbfeKeyPoint>(iKeys); F.uright = frame->uright ? S.ptr<float>(iUr) : nullptr;
    F.desc = S.ptr<uint32_t>(iDesc); F.n = n;
    F.minX = frame->min_x; F.minY = frame->min_y; F.maxX = frame->max_x; F.maxY = frame->max_y;
    F.wInv = frame->grid_w_inv; F.hInv = frame->grid_h_inv;
    F.cellStart = S.ptr<int>(wStart); F.cellItems = S.ptr<int>(wItems);
    PtsDev P;
    P.jOff = 0;
    P.m = m; P.u = S.ptr<float>(iU); P.v = S.ptr<float>(iV); P.ur = pts->ur ? S.ptr<float>(iPur) : nullptr;
    P.radius = S.ptr<float>(iRad); P.angle = pts->angle ? S.ptr<float>(iAng) : nullptr;
    P.minLevel = S.ptr<int>(iMin); P.maxLevel = S.ptr<int>(iMax); P.valid = S.ptr<uint8_t>(iVal);
    P.blocks = pts->blocks ? S.ptr<uint8_t>(iBlk) : nullptr; P.desc = S.ptr<uint32_t>(iPd);
    int* dAssigned = S.ptr<int>(ioAsg);
    int* dBestIdx = S.ptr<int>(oBest);
    int* dBestDist = S.ptr<int>(oDist);
    int* dHist = S.ptr<int>(oHist);
    int* dN = S.ptr<int>(oN);
    int* dFlag = S.ptr<int>(wFlag);

    SCK(cudaMemsetAsync(dHist, 0, 4 * (HISTO + 2), st));
    SCK(cudaMemsetAsync(dN, 0, 4, st));
    k_build_grid<<<1, 1024, 0, st>>>(F.keys, n, F.minX, F.minY, F.wInv, F.hInv, S.ptr<int>(wCellOf), S.ptr<int>(wStart),
                                     S.ptr<int>(wItems));
    const uint8_t* dcl = claimed ? S.ptr<uint8_t>(iCl) : nullptr;
    int* cin = S.ptr<int>(wA);
    int* cout = S.ptr<int>(wB);
    k_claims_init<<<(n + 255) / 256, 256, 0, st>>>(dcl, n, cin, cout);
    const int gridM = (m + 127) / 128;
    int passes = 0;
    const bool warpPass = m <= kWarpPassMax && n < (1 << 23);
    // Without conflicts the fixpoint needs two passes (the second confirms the claims of the first).  For the small
    // per-frame calls a pass costs a few microseconds and a host round trip ~20: enqueue two passes per round trip
    // (a pass on converged claims reproduces its input, so an extra one is harmless).
    const int batch = warpPass ? 2 : 1;
    for (;;) {
        // cout holds the static claims; the pass lowers entries to the first blocking acceptor
        SCK(cudaMemsetAsync(dFlag, 0, 4 * batch, st));
        for (int b = 0; b < batch; b++) {
            if (warpPass)
                k_search_pass_warp<<<(m + 3) / 4, 128, 0, st>>>(F, P, prm->mode, prm->th_accept, prm->nnratio, cin, cout, dBestIdx, dBestDist);
            else
                k_search_pass<<<gridM, 128, 0, st>>>(F, P, prm->mode, prm->th_accept, prm->nnratio, cin, cout, dBestIdx, dBestDist);
            k_claims_diff<<<(n + 255) / 256, 256, 0, st>>>(cout, cin, dcl, n, dFlag + b);
            std::swap(cin, cout);  // new claims become the input; the old table was reset by the diff
        }
        int changed[2] = {0, 0};
        SCK(cudaMemcpyAsync(changed, dFlag, 4 * batch, cudaMemcpyDeviceToHost, st));
        SCK(cudaStreamSynchronize(st));
        passes += batch;
        if (!changed[batch - 1]) break;
        if (passes > m + 2) return sfail(ORBFE_ERR_CUDA, "claim fixpoint did not converge");
    }
    k_assign_prepare<<<gridM, 128, 0, st>>>(dBestIdx, m, dAssigned);
    k_search_assign<<<gridM, 128, 0, st>>>(F, P, useHist ? 1 : 0, dBestIdx, dAssigned, dHist, S.ptr<int>(wBin), dN);
    if (useHist) k_search_cull<<<gridM, 128, 0, st>>>(m, dBestIdx, S.ptr<int>(wBin), dHist, dAssigned, dN);
    SCK(cudaGetLastError());
    SCK(S.download());
    return nmatches;
}

// SearchByProjection for a fisheye stereo frame (F.Nleft != -1), modes ORBFE_SEARCH_MAPPOINTS
// (ORBmatcher.cc:46-240 incl. :171-237) and ORBFE_SEARCH_LASTFRAME (:1951-2185 incl. :2090-2155).
extern "C" int orbfe_search_by_projection_fisheye(const OrbfeFrameView* left, const OrbfeFrameView* right,
                                                  const int32_t* l2r, const int32_t* r2l,
                                                  const OrbfeProjPoints* pl, const OrbfeProjPoints* pr,
                                                  const OrbfeSearchParams* prm, const uint8_t* claimed,
                                                  int32_t* assigned, int32_t* best_idx_left,
                                                  int32_t* best_idx_right, int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return sfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return sfail(ORBFE_ERR_INVALID, "bad device ordinal");
    SCK(cudaSetDevice(device));
    if (!left || !right || !pl || !pr || !prm || !assigned || !l2r || !r2l) return sfail(ORBFE_ERR_INVALID, "null argument");
    const int Nl = left->n, Nr = right->n, N = Nl + Nr, m = pl->m;
    if (Nl < 0 || Nr < 0 || m < 0 || pr->m != m || (prm->mode != ORBFE_SEARCH_MAPPOINTS && prm->mode != ORBFE_SEARCH_LASTFRAME))
        return sfail(ORBFE_ERR_INVALID, "bad sizes or mode (fisheye search supports modes 0 and 1)");
    for (int j = 0; j < m; j++) {
        if (best_idx_left) best_idx_left[j] = -1;
        if (best_idx_right) best_idx_right[j] = -1;
    }
    if (m == 0 || N == 0) return 0;
    const bool useHist = prm->mode == ORBFE_SEARCH_LASTFRAME && prm->check_orientation;
    if ((Nl && (!left->keys || !left->desc)) || (Nr && (!right->keys || !right->desc)) || !pl->u || !pl->v || !pl->radius ||
        !pl->min_level || !pl->max_level || !pl->desc || !pr->u || !pr->v || !pr->radius || !pr->min_level ||
        !pr->max_level || (useHist && !pl->angle))
        return sfail(ORBFE_ERR_INVALID, "missing frame / map-point array");

    OrbfeStage S;
    std::vector<uint8_t> ones(m, 1);
    const size_t M4 = 4 * (size_t)m, N4 = 4 * (size_t)N;
    const size_t iKl = S.in(left->keys, sizeof(OrbfeKeyPoint) * (size_t)Nl), iKr = S.in(right->keys, sizeof(OrbfeKeyPoint) * (size_t)Nr);
    const size_t iDl = S.in(left->desc, 32 * (size_t)Nl), iDr = S.in(right->desc, 32 * (size_t)Nr);
    const size_t iL2R = S.in(l2r, 4 * (size_t)Nl), iR2L = S.in(r2l, 4 * (size_t)Nr);
    const size_t iLu = S.in(pl->u, M4), iLv = S.in(pl->v, M4), iLr = S.in(pl->radius, M4);
    const size_t iLmin = S.in(pl->min_level, M4), iLmax = S.in(pl->max_level, M4);
    const size_t iLval = S.in(pl->valid ? pl->valid : ones.data(), (size_t)m), iLang = S.in(pl->angle, pl->angle ? M4 : 0);
    const size_t iLblk = S.in(pl->blocks, pl->blocks ? (size_t)m : 0), iLdesc = S.in(pl->desc, 32 * (size_t)m);
    const size_t iRu = S.in(pr->u, M4), iRv = S.in(pr->v, M4), iRr = S.in(pr->radius, M4);
    const size_t iRmin = S.in(pr->min_level, M4), iRmax = S.in(pr->max_level, M4);
    const size_t iRval = S.in(pr->valid ? pr->valid : ones.data(), (size_t)m), iCl = S.in(claimed, claimed ? (size_t)N : 0);
    const size_t ioAsg = S.inout(assigned, N4);
    const size_t wOf = S.work(4 * (size_t)std::max(Nl, Nr)), wStL = S.work(4 * (GC * GR + 1)), wItL = S.work(4 * (size_t)Nl);
    const size_t wStR = S.work(4 * (GC * GR + 1)), wItR = S.work(4 * (size_t)Nr), wA = S.work(N4), wB = S.work(N4);
    const size_t wTmax = S.work(N4), wHist = S.work(4 * (HISTO + 2)), wBinL = S.work(M4), wBinR = S.work(M4), wFlag = S.work(256);
    int nmatches = 0;
    const size_t oBL = S.out(best_idx_left, M4), oBR = S.out(best_idx_right, M4), oN = S.out(&nmatches, 4);
    SCK(S.commit(device));
    cudaStream_t st = S.stream();
    SCK(S.upload());
    SCK(cudaMemsetAsync(S.ptr<int>(wTmax), 0xFF, N4, st));
    SCK(cudaMemsetAsync(S.ptr<int>(wHist), 0, 4 * (HISTO + 2), st));
    SCK(cudaMemsetAsync(S.ptr<int>(oN), 0, 4, st));

    FeDev F;
    auto fill = [&](GridDev& G, const OrbfeFrameView* fv, size_t k, size_t d, size_t stt, size_t it) {
        G.keys = S.ptr<OrbfeKeyPoint>(k); G.uright = nullptr; G.desc = S.ptr<uint32_t>(d); G.n = fv->n;
        G.minX = fv->min_x; G.minY = fv->min_y; G.maxX = fv->max_x; G.maxY = fv->max_y;
        G.wInv = fv->grid_w_inv; G.hInv = fv->grid_h_inv;
        G.cellStart = S.ptr<int>(stt); G.cellItems = S.ptr<int>(it);
    };
    fill(F.L, left, iKl, iDl, wStL, wItL);
    fill(F.R, right, iKr, iDr, wStR, wItR);
    F.l2r = S.ptr<int>(iL2R); F.r2l = S.ptr<int>(iR2L); F.Nl = Nl;
    PtsFe P;
    P.L.jOff = 0;
    P.L.m = m; P.L.u = S.ptr<float>(iLu); P.L.v = S.ptr<float>(iLv); P.L.ur = nullptr; P.L.radius = S.ptr<float>(iLr);
    P.L.angle = pl->angle ? S.ptr<float>(iLang) : nullptr; P.L.minLevel = S.ptr<int>(iLmin); P.L.maxLevel = S.ptr<int>(iLmax);
    P.L.valid = S.ptr<uint8_t>(iLval); P.L.blocks = pl->blocks ? S.ptr<uint8_t>(iLblk) : nullptr; P.L.desc = S.ptr<uint32_t>(iLdesc);
    P.u = S.ptr<float>(iRu); P.v = S.ptr<float>(iRv); P.radius = S.ptr<float>(iRr); P.minLevel = S.ptr<int>(iRmin);
    P.maxLevel = S.ptr<int>(iRmax); P.valid = S.ptr<uint8_t>(iRval);
    int* dBL = S.ptr<int>(oBL);
    int* dBR = S.ptr<int>(oBR);
    int* dAssigned = S.ptr<int>(ioAsg);
    int* dN = S.ptr<int>(oN);
    int* dFlag = S.ptr<int>(wFlag);

    k_build_grid<<<1, 1024, 0, st>>>(F.L.keys, Nl, F.L.minX, F.L.minY, F.L.wInv, F.L.hInv, S.ptr<int>(wOf), S.ptr<int>(wStL), S.ptr<int>(wItL));
    k_build_grid<<<1, 1024, 0, st>>>(F.R.keys, Nr, F.R.minX, F.R.minY, F.R.wInv, F.R.hInv, S.ptr<int>(wOf), S.ptr<int>(wStR), S.ptr<int>(wItR));
    const uint8_t* dcl = claimed ? S.ptr<uint8_t>(iCl) : nullptr;
    int* cin = S.ptr<int>(wA);
    int* cout = S.ptr<int>(wB);
    k_claims_init<<<(N + 255) / 256, 256, 0, st>>>(dcl, N, cin, cout);
    const int gridM = (m + 127) / 128;

    // A non-blocking point that overwrites a slot through a stereo partner write un-blocks it: only
    // the ordered pass expresses that (never happens for local-map points, which are all observed).
    bool sequential = false;
    if (prm->mode == ORBFE_SEARCH_MAPPOINTS && pl->blocks) {
        bool anyFree = false, anyLink = false;
        for (int j = 0; j < m && !anyFree; j++) anyFree = !pl->blocks[j];
        for (int i = 0; i < Nl && !anyLink; i++) anyLink = l2r[i] != -1;
        for (int i = 0; i < Nr && !anyLink; i++) anyLink = r2l[i] != -1;
        sequential = anyFree && anyLink;
    }
    if (sequential) {
        k_search_fe_sequential<<<1, 32, 0, st>>>(F, P, prm->mode, prm->th_accept, prm->nnratio, cin, dAssigned, dBL, dBR, dN);
    } else {
        int passes = 0;
        for (;;) {
            SCK(cudaMemsetAsync(dFlag, 0, 4, st));
            k_search_pass_fe<<<gridM, 128, 0, st>>>(F, P, prm->mode, prm->th_accept, prm->nnratio, cin, cout, dBL, dBR);
            k_claims_diff<<<(N + 255) / 256, 256, 0, st>>>(cout, cin, dcl, N, dFlag);
            int changed = 0;
            SCK(cudaMemcpyAsync(&changed, dFlag, 4, cudaMemcpyDeviceToHost, st));
            SCK(cudaStreamSynchronize(st));
            passes++;
            std::swap(cin, cout);
            if (!changed) break;
            if (passes > 2 * m + 2) return sfail(ORBFE_ERR_CUDA, "claim fixpoint did not converge");
        }
        k_fe_assign<<<gridM, 128, 0, st>>>(F, P, prm->mode, useHist ? 1 : 0, dBL, dBR, S.ptr<int>(wTmax), S.ptr<int>(wHist),
                                           S.ptr<int>(wBinL), S.ptr<int>(wBinR), dN);
        k_fe_commit<<<(N + 255) / 256, 256, 0, st>>>(N, S.ptr<int>(wTmax), dAssigned);
        if (useHist)
            k_fe_cull<<<gridM, 128, 0, st>>>(m, Nl, dBL, dBR, S.ptr<int>(wBinL), S.ptr<int>(wBinR), S.ptr<int>(wHist), dAssigned, dN);
    }
    SCK(cudaGetLastError());
    SCK(S.download());
    return nmatches;
}


// int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
//                                         vector<int>& vnMatches12, int windowSize)   ORBmatcher.cc:735-891
extern "C" int orbfe_search_for_initialization(const OrbfeFrameView* f1, const OrbfeFrameView* f2, float* prev_matched,
                                               int window_size, float nnratio, int check_orientation,
                                               int32_t* matches12, int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return sfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return sfail(ORBFE_ERR_INVALID, "bad device ordinal");
    if (!f1 || !f2 || !matches12) return sfail(ORBFE_ERR_INVALID, "null argument");
    const int n1 = f1->n, n2 = f2->n;
    if (n1 < 0 || n2 < 0) return sfail(ORBFE_ERR_INVALID, "bad sizes");
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return 0;
    if (!f1->keys || !f1->desc || !f2->keys || !f2->desc || !prev_matched) return sfail(ORBFE_ERR_INVALID, "missing array");
    OrbfeStage S;
    const size_t iK1 = S.in(f1->keys, sizeof(OrbfeKeyPoint) * (size_t)n1), iD1 = S.in(f1->desc, 32 * (size_t)n1);
    const size_t iK2 = S.in(f2->keys, sizeof(OrbfeKeyPoint) * (size_t)n2), iD2 = S.in(f2->desc, 32 * (size_t)n2);
    const size_t ioPrev = S.inout(prev_matched, 8 * (size_t)n1);
    const size_t wOf = S.work(4 * (size_t)n2), wSt = S.work(4 * (GC * GR + 1)), wIt = S.work(4 * (size_t)n2);
    const size_t wM21 = S.work(4 * (size_t)n2), wMd = S.work(4 * (size_t)n2), wBin = S.work(4 * (size_t)n1);
    int nmatches = 0;
    const size_t oM12 = S.out(matches12, 4 * (size_t)n1), oN = S.out(&nmatches, 4);
    SCK(S.commit(device));
    cudaStream_t st = S.stream();
    SCK(S.upload());
    GridDev F;
    F.keys = S.ptr<OrbfeKeyPoint>(iK2); F.uright = nullptr; F.desc = S.ptr<uint32_t>(iD2); F.n = n2;
    F.minX = f2->min_x; F.minY = f2->min_y; F.maxX = f2->max_x; F.maxY = f2->max_y;
    F.wInv = f2->grid_w_inv; F.hInv = f2->grid_h_inv;
    F.cellStart = S.ptr<int>(wSt); F.cellItems = S.ptr<int>(wIt);
    k_build_grid<<<1, 1024, 0, st>>>(F.keys, n2, F.minX, F.minY, F.wInv, F.hInv, S.ptr<int>(wOf), S.ptr<int>(wSt), S.ptr<int>(wIt));
    k_search_init<<<1, 32, 0, st>>>(S.ptr<OrbfeKeyPoint>(iK1), S.ptr<uint32_t>(iD1), n1, F, S.ptr<float>(ioPrev),
                                    (float)window_size, nnratio, check_orientation, 50 /* TH_LOW */, S.ptr<int>(oM12),
                                    S.ptr<int>(wM21), S.ptr<int>(wMd), S.ptr<int>(wBin), S.ptr<int>(oN));
    SCK(cudaGetLastError());
    SCK(S.download());
    return nmatches;
}

namespace {

// Staging of one keyframe view + one point set for the stateless window searches.
struct WinStage {
    size_t iKeys, iDesc, iUr, iU, iV, iPur, iRad, iMin, iMax, iVal, iPd, wCellOf, wStart, wItems;
    std::vector<uint8_t> ones;
    void lay(OrbfeStage& S, const OrbfeFrameView* f, const OrbfeProjPoints* p) {
        const size_t n = (size_t)f->n, m = (size_t)p->m;
        if (!p->valid) ones.assign(m, 1);
        iKeys = S.in(f->keys, sizeof(OrbfeKeyPoint) * n); iDesc = S.in(f->desc, 32 * n);
        iUr = S.in(f->uright, f->uright ? 4 * n : 0);
        iU = S.in(p->u, 4 * m); iV = S.in(p->v, 4 * m); iPur = S.in(p->ur, p->ur ? 4 * m : 0);
        iRad = S.in(p->radius, 4 * m); iMin = S.in(p->min_level, 4 * m); iMax = S.in(p->max_level, 4 * m);
        iVal = S.in(p->valid ? p->valid : ones.data(), m); iPd = S.in(p->desc, 32 * m);
        wCellOf = S.work(4 * n); wStart = S.work(4 * (GC * GR + 1)); wItems = S.work(4 * n);
    }
    void bind(const OrbfeStage& S, const OrbfeFrameView* f, const OrbfeProjPoints* p, GridDev& F, PtsDev& P) const {
        F.keys = S.ptr<OrbfeKeyPoint>(iKeys); F.uright = f->uright ? S.ptr<float>(iUr) : nullptr;
        F.desc = S.ptr<uint32_t>(iDesc); F.n = f->n;
        F.minX = f->min_x; F.minY = f->min_y; F.maxX = f->max_x; F.maxY = f->max_y;
        F.wInv = f->grid_w_inv; F.hInv = f->grid_h_inv;
        F.cellStart = S.ptr<int>(wStart); F.cellItems = S.ptr<int>(wItems);
        P.m = p->m; P.u = S.ptr<float>(iU); P.v = S.ptr<float>(iV); P.ur = p->ur ? S.ptr<float>(iPur) : nullptr;
        P.radius = S.ptr<float>(iRad); P.angle = nullptr;
        P.minLevel = S.ptr<int>(iMin); P.maxLevel = S.ptr<int>(iMax); P.valid = S.ptr<uint8_t>(iVal);
        P.blocks = nullptr; P.desc = S.ptr<uint32_t>(iPd);
    }
    void grid(const OrbfeStage& S, const GridDev& F, cudaStream_t st) const {
        k_build_grid<<<1, 1024, 0, st>>>(F.keys, F.n, F.minX, F.minY, F.wInv, F.hInv, S.ptr<int>(wCellOf),
                                         S.ptr<int>(wStart), S.ptr<int>(wItems));
    }
};

int check_window_args(const OrbfeFrameView* f, const OrbfeProjPoints* p, int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return sfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return sfail(ORBFE_ERR_INVALID, "bad device ordinal");
    if (!f || !p) return sfail(ORBFE_ERR_INVALID, "null argument");
    if (f->n < 0 || p->m < 0 || f->n >= (1 << 23)) return sfail(ORBFE_ERR_INVALID, "bad sizes");
    if (f->n > 0 && p->m > 0 &&
        (!f->keys || !f->desc || !p->u || !p->v || !p->radius || !p->min_level || !p->max_level || !p->desc))
        return sfail(ORBFE_ERR_INVALID, "missing keyframe / map-point array");
    return ORBFE_OK;
}

}  // namespace

extern "C" int orbfe_search_window(const OrbfeFrameView* kf, const OrbfeProjPoints* pts, const OrbfeWindowParams* prm,
                                   int32_t* best_idx, int32_t* best_dist, int device) {
    if (!prm || !best_idx) return sfail(ORBFE_ERR_INVALID, "null argument");
    int rc = check_window_args(kf, pts, device);
    if (rc != ORBFE_OK) return rc;
    const int m = pts->m;
    if (prm->gate != ORBFE_GATE_NONE && prm->gate != ORBFE_GATE_FUSE) return sfail(ORBFE_ERR_INVALID, "bad gate");
    if (prm->gate == ORBFE_GATE_FUSE && (!prm->inv_level_sigma2 || prm->n_levels <= 0))
        return sfail(ORBFE_ERR_INVALID, "the Fuse gate needs mvInvLevelSigma2");
    if (m == 0 || kf->n == 0) {
        for (int j = 0; j < m; j++) {
            best_idx[j] = -1;
            if (best_dist) best_dist[j] = 256;
        }
        return 0;
    }
    OrbfeStage S;
    WinStage W;
    W.lay(S, kf, pts);
    const bool gate = prm->gate == ORBFE_GATE_FUSE;
    const size_t iInv = S.in(prm->inv_level_sigma2, gate ? 4 * (size_t)prm->n_levels : 0);
    const size_t oBest = S.out(best_idx, 4 * (size_t)m), oDist = S.out(best_dist, 4 * (size_t)m);
    SCK(S.commit(device));
    cudaStream_t st = S.stream();
    SCK(S.upload());
    GridDev F;
    PtsDev P;
    P.jOff = 0;
    W.bind(S, kf, pts, F, P);
    W.grid(S, F, st);
    k_window_best<<<(m + 3) / 4, 128, 0, st>>>(F, P, prm->th_accept, gate ? 1 : 0, gate ? S.ptr<float>(iInv) : nullptr,
                                               prm->n_levels, S.ptr<int>(oBest), S.ptr<int>(oDist));
    SCK(cudaGetLastError());
    SCK(S.download());
    int n = 0;
    for (int j = 0; j < m; j++) n += best_idx[j] >= 0;
    return n;
}

extern "C" int orbfe_search_by_sim3(const OrbfeFrameView* kf1, const OrbfeFrameView* kf2, const OrbfeProjPoints* pts12,
                                    const OrbfeProjPoints* pts21, int th_accept, int32_t* match12, int device) {
    if (!match12) return sfail(ORBFE_ERR_INVALID, "null argument");
    int rc = check_window_args(kf2, pts12, device);
    if (rc == ORBFE_OK) rc = check_window_args(kf1, pts21, device);
    if (rc != ORBFE_OK) return rc;
    if (pts12->m != kf1->n || pts21->m != kf2->n)
        return sfail(ORBFE_ERR_INVALID, "SearchBySim3 takes one projected point per keyframe slot");
    const int n1 = kf1->n, n2 = kf2->n;
    if (n1 == 0 || n2 == 0) {
        for (int i = 0; i < n1; i++) match12[i] = -1;
        return 0;
    }
    OrbfeStage S;
    WinStage W12, W21;   // W12: KF1's points searched in KF2; W21: the reverse
    W12.lay(S, kf2, pts12);
    W21.lay(S, kf1, pts21);
    const size_t wM1 = S.work(4 * (size_t)n1), wM2 = S.work(4 * (size_t)n2), wD = S.work(4 * (size_t)std::max(n1, n2));
    int nFound = 0;
    const size_t oMatch = S.out(match12, 4 * (size_t)n1), oN = S.out(&nFound, 4);
    SCK(S.commit(device));
    cudaStream_t st = S.stream();
    SCK(S.upload());
    GridDev F1, F2;
    PtsDev P12, P21;
    W12.bind(S, kf2, pts12, F2, P12);
    W21.bind(S, kf1, pts21, F1, P21);
    W12.grid(S, F2, st);
    W21.grid(S, F1, st);
    SCK(cudaMemsetAsync(S.ptr<int>(oN), 0, 4, st));
    k_window_best<<<(n1 + 3) / 4, 128, 0, st>>>(F2, P12, th_accept, 0, nullptr, 0, S.ptr<int>(wM1), S.ptr<int>(wD));
    k_window_best<<<(n2 + 3) / 4, 128, 0, st>>>(F1, P21, th_accept, 0, nullptr, 0, S.ptr<int>(wM2), S.ptr<int>(wD));
    k_sim3_mutual<<<(n1 + 255) / 256, 256, 0, st>>>(n1, S.ptr<int>(wM1), S.ptr<int>(wM2), S.ptr<int>(oMatch), S.ptr<int>(oN));
    SCK(cudaGetLastError());
    SCK(S.download());
    return nFound;
}

// ---------------------------------------------------------------------------------------------
// Map shards: SearchByProjection(Frame&, vector<MapPoint*>&, ...) (ORBmatcher.cc:46-240) against a map that is split
// over several GPUs by contiguous index ranges (BASELINE config 5).  The reference's loop is sequential over the map
// points (:54) and a keypoint accepted by point j is skipped by every later point (:103-105); the claim fixpoint above
// keeps that exact: a claim is the GLOBAL index of the first accepting point, every shard runs the pass over its own
// points against the same global claim table, the shards' new tables are combined by an elementwise minimum (the only
// exchange: 4 bytes per frame keypoint per pass), and the passes repeat until the table stops changing -- then every
// point has seen exactly the claims of the points before it, on whichever GPU they live.
struct OrbfeMapShard {
    int device = 0, m = 0, j0 = 0, n = 0;
    bool hasUr = false, hasBlocks = false, hasAngle = false, hasFrameUr = false;
    char* dPts = nullptr;     // u, v, ur, radius, minLevel, maxLevel, valid, blocks, desc
    char* dFrame = nullptr;   // keys, desc, uright, cellOf, cellStart, cellItems
    size_t frameCap = 0;
    PtsDev P;
    GridDev F;
    int *dBestIdx = nullptr, *dBestDist = nullptr;
};

namespace {
struct PeerClaims { const int32_t* t[16]; };
// Elementwise minimum of the shards' claim tables, read with peer loads (symmetric memory over NVLink): the exchange
// step of the sharded search fused into the kernel that produces the next pass's input.
__global__ void k_claims_min_peers(PeerClaims C, int G, int n, int32_t* __restrict__ out, const int32_t* __restrict__ prev,
                                   int* __restrict__ changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = INT_MAX;
    for (int s = 0; s < G; s++) v = min(v, C.t[s][i]);
    out[i] = v;
    if (prev && prev[i] != v) *changed = 1;
}
size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace

extern "C" int orbfe_map_shard_create(const OrbfeProjPoints* pts, int j0, int device, OrbfeMapShard** out) {
    if (!out) return sfail(ORBFE_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return sfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return sfail(ORBFE_ERR_INVALID, "bad device ordinal");
    if (!pts || pts->m <= 0 || j0 < 0 || !pts->u || !pts->v || !pts->radius || !pts->min_level || !pts->max_level || !pts->desc)
        return sfail(ORBFE_ERR_INVALID, "missing map-point array");
    SCK(cudaSetDevice(device));
    const size_t m = (size_t)pts->m, M4 = 4 * m;
    OrbfeMapShard* h = new OrbfeMapShard();
    h->device = device; h->m = pts->m; h->j0 = j0;
    h->hasUr = pts->ur != nullptr; h->hasBlocks = pts->blocks != nullptr; h->hasAngle = pts->angle != nullptr;
    size_t off[10], total = 0;
    const size_t sz[10] = {M4, M4, M4, M4, M4, M4, m, m, 32 * m, M4};
    for (int i = 0; i < 10; i++) { off[i] = total; total += al256(sz[i]); }
    cudaError_t e = cudaMalloc(&h->dPts, total + 2 * al256(M4));
    if (e != cudaSuccess) { delete h; return sfail(ORBFE_ERR_CUDA, "cudaMalloc(map shard)", e); }
    char* d = h->dPts;
    std::vector<uint8_t> ones;
    if (!pts->valid) ones.assign(m, 1);
    const void* src[10] = {pts->u, pts->v, pts->ur, pts->radius, pts->min_level, pts->max_level,
                           pts->valid ? pts->valid : ones.data(), pts->blocks, pts->desc, pts->angle};
    for (int i = 0; i < 10 && e == cudaSuccess; i++)
        if (src[i]) e = cudaMemcpy(d + off[i], src[i], sz[i], cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(h->dPts); delete h; return sfail(ORBFE_ERR_CUDA, "upload of the map shard", e); }
    PtsDev& P = h->P;
    P.m = pts->m; P.jOff = j0;
    P.u = (const float*)(d + off[0]); P.v = (const float*)(d + off[1]); P.ur = pts->ur ? (const float*)(d + off[2]) : nullptr;
    P.radius = (const float*)(d + off[3]); P.minLevel = (const int*)(d + off[4]); P.maxLevel = (const int*)(d + off[5]);
    P.valid = (const uint8_t*)(d + off[6]); P.blocks = pts->blocks ? (const uint8_t*)(d + off[7]) : nullptr;
    P.desc = (const uint32_t*)(d + off[8]); P.angle = pts->angle ? (const float*)(d + off[9]) : nullptr;
    h->dBestIdx = (int*)(d + total);
    h->dBestDist = (int*)(d + total + al256(M4));
    *out = h;
    return ORBFE_OK;
}

extern "C" void orbfe_map_shard_destroy(OrbfeMapShard* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->dPts) cudaFree(h->dPts);
    if (h->dFrame) cudaFree(h->dFrame);
    delete h;
}

// The frame side, the same on every shard: keypoints, descriptors, stereo coordinates; builds the 64 x 48 grid.
extern "C" int orbfe_map_shard_set_frame(OrbfeMapShard* h, const OrbfeFrameView* frame, void* stream) {
    if (!h || !frame || frame->n <= 0 || !frame->keys || !frame->desc) return sfail(ORBFE_ERR_INVALID, "bad frame view");
    SCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)frame->n, N4 = 4 * n;
    const size_t sz[6] = {sizeof(OrbfeKeyPoint) * n, 32 * n, N4, N4, 4 * (size_t)(GC * GR + 1), N4};
    size_t off[6], total = 0;
    for (int i = 0; i < 6; i++) { off[i] = total; total += al256(sz[i]); }
    if (total > h->frameCap) {
        SCK(cudaStreamSynchronize(st));
        if (h->dFrame) cudaFree(h->dFrame);
        h->dFrame = nullptr; h->frameCap = 0;
        SCK(cudaMalloc(&h->dFrame, total));
        h->frameCap = total;
    }
    char* d = h->dFrame;
    SCK(cudaMemcpyAsync(d + off[0], frame->keys, sz[0], cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d + off[1], frame->desc, sz[1], cudaMemcpyHostToDevice, st));
    if (frame->uright) SCK(cudaMemcpyAsync(d + off[2], frame->uright, sz[2], cudaMemcpyHostToDevice, st));
    GridDev& F = h->F;
    F.keys = (const OrbfeKeyPoint*)(d + off[0]); F.desc = (const uint32_t*)(d + off[1]);
    F.uright = frame->uright ? (const float*)(d + off[2]) : nullptr;
    F.n = frame->n; h->n = frame->n;
    F.minX = frame->min_x; F.minY = frame->min_y; F.maxX = frame->max_x; F.maxY = frame->max_y;
    F.wInv = frame->grid_w_inv; F.hInv = frame->grid_h_inv;
    F.cellStart = (const int*)(d + off[4]); F.cellItems = (const int*)(d + off[5]);
    k_build_grid<<<1, 1024, 0, st>>>(F.keys, F.n, F.minX, F.minY, F.wInv, F.hInv, (int*)(d + off[3]), (int*)(d + off[4]), (int*)(d + off[5]));
    SCK(cudaGetLastError());
    return ORBFE_OK;
}

// Static claims of a frame: -1 where the keypoint already holds a blocking map point, INT_MAX elsewhere (device arrays).
extern "C" int orbfe_claims_init_device(const uint8_t* d_claimed, int n, int32_t* d_claims, void* stream) {
    if (n <= 0) return ORBFE_OK;
    if (!d_claims) return sfail(ORBFE_ERR_INVALID, "null claims table");
    k_claims_init<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_claimed, n, d_claims, d_claims);
    SCK(cudaGetLastError());
    return ORBFE_OK;
}

// One pass of the shard's points (global indices j0 .. j0+m) against the global claim table d_claim_in; d_claim_out holds
// the static claims on entry and is lowered to the first accepting point of this shard.
extern "C" int orbfe_map_shard_pass(OrbfeMapShard* h, const OrbfeSearchParams* prm, const int32_t* d_claim_in,
                                    int32_t* d_claim_out, void* stream) {
    if (!h || !prm || !d_claim_in || !d_claim_out || !h->dFrame) return sfail(ORBFE_ERR_INVALID, "map shard: frame not set / null argument");
    if (prm->mode != ORBFE_SEARCH_MAPPOINTS && !(prm->mode == ORBFE_SEARCH_KEYFRAME && !prm->check_orientation))
        return sfail(ORBFE_ERR_INVALID, "map shards support SearchByProjection(Frame, map points) and the keyframe mode without orientation check");
    SCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int m = h->m;
    if (m <= kWarpPassMax && h->n < (1 << 23))
        k_search_pass_warp<<<(m + 3) / 4, 128, 0, st>>>(h->F, h->P, prm->mode, prm->th_accept, prm->nnratio, d_claim_in, d_claim_out,
                                                       h->dBestIdx, h->dBestDist);
    else
        k_search_pass<<<(m + 127) / 128, 128, 0, st>>>(h->F, h->P, prm->mode, prm->th_accept, prm->nnratio, d_claim_in, d_claim_out,
                                                     h->dBestIdx, h->dBestDist);
    SCK(cudaGetLastError());
    return ORBFE_OK;
}

// The exchange step over peer memory: d_out[i] = min over the shards' claim tables (peer pointers, e.g. symmetric
// memory over NVLink); *d_changed is set when the result differs from d_prev (the table the pass just read).
extern "C" int orbfe_claims_min_peers_device(const int32_t* const* peer_tabs, int G, int n, int32_t* d_out, const int32_t* d_prev,
                                             int32_t* d_changed, void* stream) {
    if (n <= 0 || G <= 0) return ORBFE_OK;
    if (!peer_tabs || !d_out || G > 16) return sfail(ORBFE_ERR_INVALID, "claims exchange: null argument or more than 16 shards");
    PeerClaims C;
    for (int s = 0; s < 16; s++) C.t[s] = s < G ? peer_tabs[s] : nullptr;
    k_claims_min_peers<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(C, G, n, d_out, d_prev, d_changed);
    SCK(cudaGetLastError());
    return ORBFE_OK;
}

// After the fixpoint: raises d_assigned[k] (global point indices; -1 / the caller's previous content elsewhere) for the
// keypoints this shard's points accepted and adds the shard's match count to *d_nmatches.  The per-shard tables are
// combined by an elementwise maximum (the reference's later point overwrites an earlier one, :156) and a sum.
extern "C" int orbfe_map_shard_finish(OrbfeMapShard* h, int32_t* d_assigned, int32_t* d_nmatches, void* stream) {
    if (!h || !d_assigned || !d_nmatches || !h->dFrame) return sfail(ORBFE_ERR_INVALID, "map shard: null argument");
    SCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int gridM = (h->m + 127) / 128;
    k_search_assign<<<gridM, 128, 0, st>>>(h->F, h->P, 0, h->dBestIdx, d_assigned, nullptr, nullptr, d_nmatches);
    SCK(cudaGetLastError());
    return ORBFE_OK;
}

// Per-point results of the last pass (host arrays of the shard's m points; indices are frame keypoint indices).
extern "C" int orbfe_map_shard_results(OrbfeMapShard* h, int32_t* best_idx, int32_t* best_dist, void* stream) {
    if (!h) return sfail(ORBFE_ERR_INVALID, "null map shard");
    SCK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (best_idx) SCK(cudaMemcpyAsync(best_idx, h->dBestIdx, 4 * (size_t)h->m, cudaMemcpyDeviceToHost, st));
    if (best_dist) SCK(cudaMemcpyAsync(best_dist, h->dBestDist, 4 * (size_t)h->m, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    return ORBFE_OK;
}
