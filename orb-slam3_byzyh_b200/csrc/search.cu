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

namespace {

constexpr int GC = 64, GR = 48;          // FRAME_GRID_COLS / ROWS, include/Frame.h:44-45
constexpr int HISTO = 30;                // HISTO_LENGTH, ORBmatcher.cc:38

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
                        if (claimIn[idx] < j) continue;  // ORBmatcher.cc:103-105 / 2040-2042 / 2266-2267
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
    if (accept && (mode == ORBFE_SEARCH_KEYFRAME || (P.blocks ? P.blocks[j] != 0 : true))) atomicMin(&claimOut[bI], j);
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
    atomicMax(&assigned[k], j);
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

int sfail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define SCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return sfail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
    cudaError_t upload(const void* src, size_t n) {
        cudaError_t e = alloc(n);
        if (e == cudaSuccess && n) e = cudaMemcpy(p, src, n, cudaMemcpyHostToDevice);
        return e;
    }
    template <class T> T* as() { return (T*)p; }
};

}  // namespace

extern "C" int orbfe_search_by_projection(const OrbfeFrameView* frame, const OrbfeProjPoints* pts,
                                          const OrbfeSearchParams* prm, const uint8_t* claimed,
                                          int32_t* assigned, int32_t* best_idx, int32_t* best_dist, int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return sfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return sfail(ORBFE_ERR_INVALID, "bad device ordinal");
    SCK(cudaSetDevice(device));
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

    DevBuf dKeys, dUr, dDesc, dCellOf, dCellStart, dCellItems;
    SCK(dKeys.upload(frame->keys, sizeof(OrbfeKeyPoint) * (size_t)n));
    SCK(dDesc.upload(frame->desc, 32 * (size_t)n));
    if (frame->uright) SCK(dUr.upload(frame->uright, 4 * (size_t)n));
    SCK(dCellOf.alloc(4 * (size_t)n)); SCK(dCellStart.alloc(4 * (GC * GR + 1))); SCK(dCellItems.alloc(4 * (size_t)n));
    DevBuf pu, pv, pur, prad, pang, pminl, pmaxl, pvalid, pblocks, pdesc;
    SCK(pu.upload(pts->u, 4 * (size_t)m)); SCK(pv.upload(pts->v, 4 * (size_t)m));
    if (pts->ur) SCK(pur.upload(pts->ur, 4 * (size_t)m));
    SCK(prad.upload(pts->radius, 4 * (size_t)m));
    if (pts->angle) SCK(pang.upload(pts->angle, 4 * (size_t)m));
    SCK(pminl.upload(pts->min_level, 4 * (size_t)m)); SCK(pmaxl.upload(pts->max_level, 4 * (size_t)m));
    std::vector<uint8_t> ones;
    if (!pts->valid) ones.assign(m, 1);
    SCK(pvalid.upload(pts->valid ? pts->valid : ones.data(), (size_t)m));
    if (pts->blocks) SCK(pblocks.upload(pts->blocks, (size_t)m));
    SCK(pdesc.upload(pts->desc, 32 * (size_t)m));
    DevBuf dClaimed, cA, cB, dBestIdx, dBestDist, dAssigned, dHist, dBin, dFlag;
    if (claimed) SCK(dClaimed.upload(claimed, (size_t)n));
    SCK(cA.alloc(4 * (size_t)n)); SCK(cB.alloc(4 * (size_t)n));
    SCK(dBestIdx.alloc(4 * (size_t)m)); SCK(dBestDist.alloc(4 * (size_t)m));
    SCK(dAssigned.upload(assigned, 4 * (size_t)n));
    SCK(dHist.alloc(4 * (HISTO + 2))); SCK(dBin.alloc(4 * (size_t)m)); SCK(dFlag.alloc(4));
    SCK(cudaMemset(dHist.p, 0, 4 * (HISTO + 2)));

    GridDev F;
    F.keys = dKeys.as<OrbfeKeyPoint>(); F.uright = frame->uright ? dUr.as<float>() : nullptr;
    F.desc = dDesc.as<uint32_t>(); F.n = n;
    F.minX = frame->min_x; F.minY = frame->min_y; F.maxX = frame->max_x; F.maxY = frame->max_y;
    F.wInv = frame->grid_w_inv; F.hInv = frame->grid_h_inv;
    F.cellStart = dCellStart.as<int>(); F.cellItems = dCellItems.as<int>();
    PtsDev P;
    P.m = m; P.u = pu.as<float>(); P.v = pv.as<float>(); P.ur = pts->ur ? pur.as<float>() : nullptr;
    P.radius = prad.as<float>(); P.angle = pts->angle ? pang.as<float>() : nullptr;
    P.minLevel = pminl.as<int>(); P.maxLevel = pmaxl.as<int>(); P.valid = pvalid.as<uint8_t>();
    P.blocks = pts->blocks ? pblocks.as<uint8_t>() : nullptr; P.desc = pdesc.as<uint32_t>();

    k_build_grid<<<1, 1024>>>(F.keys, n, F.minX, F.minY, F.wInv, F.hInv, dCellOf.as<int>(), dCellStart.as<int>(),
                              dCellItems.as<int>());
    const uint8_t* dcl = claimed ? dClaimed.as<uint8_t>() : nullptr;
    k_claims_init<<<(n + 255) / 256, 256>>>(dcl, n, cA.as<int>(), cB.as<int>());
    int* cin = cA.as<int>();
    int* cout = cB.as<int>();
    const int gridM = (m + 127) / 128;
    int passes = 0;
    for (;;) {
        // cout holds the static claims; the pass lowers entries to the first blocking acceptor
        k_search_pass<<<gridM, 128>>>(F, P, prm->mode, prm->th_accept, prm->nnratio, cin, cout,
                                      dBestIdx.as<int>(), dBestDist.as<int>());
        SCK(cudaMemset(dFlag.p, 0, 4));
        k_claims_diff<<<(n + 255) / 256, 256>>>(cout, cin, dcl, n, dFlag.as<int>());
        int changed = 0;
        SCK(cudaMemcpy(&changed, dFlag.p, 4, cudaMemcpyDeviceToHost));
        passes++;
        std::swap(cin, cout);  // new claims become the input; the old table was reset by the diff
        if (!changed) break;
        if (passes > m + 1) return sfail(ORBFE_ERR_CUDA, "claim fixpoint did not converge");
    }
    int* dN = dHist.as<int>() + HISTO;
    k_assign_prepare<<<gridM, 128>>>(dBestIdx.as<int>(), m, dAssigned.as<int>());
    k_search_assign<<<gridM, 128>>>(F, P, useHist ? 1 : 0, dBestIdx.as<int>(), dAssigned.as<int>(), dHist.as<int>(),
                                    dBin.as<int>(), dN);
    if (useHist) k_search_cull<<<gridM, 128>>>(m, dBestIdx.as<int>(), dBin.as<int>(), dHist.as<int>(), dAssigned.as<int>(), dN);
    SCK(cudaGetLastError());
    int nmatches = 0;
    SCK(cudaMemcpy(&nmatches, dN, 4, cudaMemcpyDeviceToHost));
    SCK(cudaMemcpy(assigned, dAssigned.p, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    if (best_idx) SCK(cudaMemcpy(best_idx, dBestIdx.p, 4 * (size_t)m, cudaMemcpyDeviceToHost));
    if (best_dist) SCK(cudaMemcpy(best_dist, dBestDist.p, 4 * (size_t)m, cudaMemcpyDeviceToHost));
    return nmatches;
}
