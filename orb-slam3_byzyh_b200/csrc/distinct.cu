// distinct.cu -- void MapPoint::ComputeDistinctiveDescriptors() (/root/reference/src/MapPoint.cc:438-529), batched over
// map points (LocalMapping calls it for every new / fused point, src/LocalMapping.cc:430, 975-1040).
// Per point with N observed descriptors: all-pairs Hamming distances, per row the median = sorted[(N-1)/2] with the
// row's own 0 included, best = first row with the smallest median (strict <).  One warp per map point; the median of a
// row needs no sort: distances are integers in [0, 256], so a 257-bin histogram in shared memory and one warp scan give
// the order statistic exactly.
#include "orbfe_internal.h"
#include "scratch.h"

namespace {

int dfail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define DCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return dfail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

constexpr int DW = 4;   // warps (map points) per CTA

__global__ void __launch_bounds__(32 * DW)
k_distinctive(const uint32_t* __restrict__ desc, const int* __restrict__ start, int nPoints, int* __restrict__ bestIdx) {
    __shared__ int hist[DW][264];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * DW + w;
    if (p >= nPoints) return;
    const int b = start[p], N = start[p + 1] - b;
    if (N <= 0) {
        if (lane == 0) bestIdx[p] = -1;
        return;
    }
    int* h = hist[w];
    const int rank = (N - 1) >> 1;   // vDists[0.5*(N-1)]
    int bestMedian = 0x7fffffff, best = 0;
    for (int i = 0; i < N; i++) {
        for (int k = lane; k < 264; k += 32) h[k] = 0;
        __syncwarp();
        const uint4* pi = reinterpret_cast<const uint4*>(desc + 8 * (size_t)(b + i));
        const uint4 a0 = pi[0], a1 = pi[1];
        for (int j = lane; j < N; j += 32) {
            const uint4* pj = reinterpret_cast<const uint4*>(desc + 8 * (size_t)(b + j));
            const uint4 c0 = pj[0], c1 = pj[1];
            const int d = __popc(a0.x ^ c0.x) + __popc(a0.y ^ c0.y) + __popc(a0.z ^ c0.z) + __popc(a0.w ^ c0.w) +
                          __popc(a1.x ^ c1.x) + __popc(a1.y ^ c1.y) + __popc(a1.z ^ c1.z) + __popc(a1.w ^ c1.w);
            atomicAdd(&h[d], 1);
        }
        __syncwarp();
        // lane l owns bins 8l .. 8l+7 (+ bin 256 for lane 31, padded bins are zero)
        int c[9], s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { c[k] = h[8 * lane + k]; s += c[k]; }
        c[8] = lane == 31 ? h[256] : 0;
        s += c[8];
        int incl = s;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int cum = incl - s, median = -1;   // values before this lane's bins
#pragma unroll
        for (int k = 0; k < 9; k++) {
            if (median < 0 && cum + c[k] > rank) median = (k < 8) ? 8 * lane + k : 256;
            cum += c[k];
        }
        const unsigned found = __ballot_sync(0xffffffffu, median >= 0);
        median = __shfl_sync(0xffffffffu, median, __ffs(found) - 1);
        if (median < bestMedian) { bestMedian = median; best = i; }
        __syncwarp();
    }
    if (lane == 0) bestIdx[p] = best;
}

}  // namespace

extern "C" int orbfe_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int n_points, int32_t* best_idx,
                                             int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return dfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return dfail(ORBFE_ERR_INVALID, "bad device ordinal");
    if (n_points < 0 || (n_points > 0 && (!start || !best_idx))) return dfail(ORBFE_ERR_INVALID, "null argument");
    if (n_points == 0) return ORBFE_OK;
    const int total = start[n_points];
    if (start[0] != 0 || total < 0 || (total > 0 && !desc)) return dfail(ORBFE_ERR_INVALID, "bad descriptor ranges");
    for (int p = 0; p < n_points; p++)
        if (start[p + 1] < start[p]) return dfail(ORBFE_ERR_INVALID, "descriptor ranges must be ascending");
    OrbfeStage S;
    const size_t iD = S.in(desc, 32 * (size_t)total), iS = S.in(start, 4 * (size_t)(n_points + 1));
    const size_t oB = S.out(best_idx, 4 * (size_t)n_points);
    DCK(S.commit(device));
    DCK(S.upload());
    k_distinctive<<<(n_points + DW - 1) / DW, 32 * DW, 0, S.stream()>>>(S.ptr<uint32_t>(iD), S.ptr<int>(iS), n_points, S.ptr<int>(oB));
    DCK(cudaGetLastError());
    DCK(S.download());
    return ORBFE_OK;
}
