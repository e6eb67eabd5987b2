// match.cu -- Hamming matching kernels and their C ABI (include/orbfe.h):
//   ORBmatcher::DescriptorDistance          /root/reference/src/ORBmatcher.cc:2384-2404
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) + 0.7 ratio, Frame::ComputeStereoFishEyeMatches
//                                            src/Frame.cc:47, 1553, 1562
// 256-bit descriptors are 8 x 32-bit words: XOR (LOP3) + POPC on the integer pipes.  Best and
// second best are tracked as packed keys (distance << 23 | train index): the minimum of such
// keys is "smallest distance, ties -> lowest train index", which is BFMatcher's order and the
// reference's strict-`<` scan order, so top-2 is three VIMNMX per pair and merges (across train
// chunks, and across GPUs when the map is sharded) are associative.
#include <stdio.h>

#include <algorithm>
#include <string>
#include <vector>

#include "orbfe_internal.h"
#include "scratch.h"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int KEY_SHIFT = 23;                   // train indices < 2^23 per launch
constexpr int KNN_THREADS = 128, KNN_QPT = 2;   // queries per thread
constexpr int KNN_QB = KNN_THREADS * KNN_QPT;   // queries per CTA
constexpr int KNN_TILE = 256;                   // train descriptors staged per iteration
constexpr int KNN_CHUNK_MAX = 4096;             // train descriptors per CTA (upper bound)

__device__ __forceinline__ int hamming256(const uint32_t* a, const uint32_t* b) {
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
    return d;
}

// Same distance with half the POPCs: POPC issues at a quarter of the LOP3 rate on B200 (measured
// ~16 lanes/clk/SM, tools/pipe_bench2.cu), so the eight XOR words are first compressed with a
// Harley-Seal carry-save adder tree (LOP3 full adders) into four words of weight 1, 2, 4, 8.
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}
__device__ __forceinline__ int hamming256_csa(const uint32_t* a, const uint32_t* b) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a[i] ^ b[i];
    // three full adders: 8 words -> 2 of weight 1 (s3, x7) + 3 of weight 2 (c1, c2, c3): 5 POPC + 14 LOP3
    // balances the POPC pipe against the issue rate (a full tree down to 4 POPC costs more LOP3 than it saves)
    uint32_t s1, c1, s2, c2, s3, c3;
    csa(x[0], x[1], x[2], s1, c1);
    csa(x[3], x[4], x[5], s2, c2);
    csa(x[6], s1, s2, s3, c3);
    return (__popc(s3) + __popc(x[7])) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

__device__ __forceinline__ void top2_insert(uint32_t& b0, uint32_t& b1, uint32_t k) {
    const uint32_t hi = max(b0, k);
    b0 = min(b0, k);
    b1 = min(b1, hi);
}

__global__ void k_hamming_pairs(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int n,
                                int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x[8], y[8];
    const uint4* pa = reinterpret_cast<const uint4*>(a + 8 * (size_t)i);
    const uint4* pb = reinterpret_cast<const uint4*>(b + 8 * (size_t)i);
    *reinterpret_cast<uint4*>(x) = pa[0]; *reinterpret_cast<uint4*>(x + 4) = pa[1];
    *reinterpret_cast<uint4*>(y) = pb[0]; *reinterpret_cast<uint4*>(y + 4) = pb[1];
    out[i] = hamming256(x, y);
}

// partial[q][chunk][2] = two smallest keys of query q over train chunk `chunk`
__global__ void __launch_bounds__(KNN_THREADS)
k_knn2_partial(const uint32_t* __restrict__ query, int nq, const uint32_t* __restrict__ train, int nt,
               int chunk, int nchunks, uint32_t* __restrict__ partial) {
    __shared__ __align__(16) uint32_t tile[KNN_TILE * 8];
    const int q0 = blockIdx.x * KNN_QB + threadIdx.x;
    uint32_t qv[KNN_QPT][8];
    uint32_t b0[KNN_QPT], b1[KNN_QPT];
#pragma unroll
    for (int r = 0; r < KNN_QPT; r++) {
        const int q = min(q0 + r * KNN_THREADS, nq - 1);
        const uint4* p = reinterpret_cast<const uint4*>(query + 8 * (size_t)q);
        *reinterpret_cast<uint4*>(qv[r]) = p[0];
        *reinterpret_cast<uint4*>(qv[r] + 4) = p[1];
        b0[r] = KEY_NONE;
        b1[r] = KEY_NONE;
    }
    const int c0 = blockIdx.y * chunk, c1 = min(c0 + chunk, nt);
    for (int t0 = c0; t0 < c1; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, c1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += KNN_THREADS)
            reinterpret_cast<uint4*>(tile)[i] = reinterpret_cast<const uint4*>(train + 8 * (size_t)t0)[i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            uint32_t tv[8];
            *reinterpret_cast<uint4*>(tv) = reinterpret_cast<const uint4*>(tile)[2 * j];
            *reinterpret_cast<uint4*>(tv + 4) = reinterpret_cast<const uint4*>(tile)[2 * j + 1];
            const uint32_t jj = (uint32_t)(t0 + j);
#pragma unroll
            for (int r = 0; r < KNN_QPT; r++) {
                const uint32_t key = ((uint32_t)hamming256_csa(qv[r], tv) << KEY_SHIFT) | jj;
                top2_insert(b0[r], b1[r], key);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KNN_QPT; r++) {
        const int q = q0 + r * KNN_THREADS;
        if (q < nq) {
            uint32_t* o = partial + ((size_t)q * nchunks + blockIdx.y) * 2;
            o[0] = b0[r];
            o[1] = b1[r];
        }
    }
}

// One warp per query: merge `nparts` (key0,key1) pairs into idx2/dist2 (+ratio-tested match).
__global__ void __launch_bounds__(256)
k_knn2_merge_keys(const uint32_t* __restrict__ partial, int nq, int nparts, int train_offset,
                  int32_t* __restrict__ idx2, int32_t* __restrict__ dist2, int32_t* __restrict__ match) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t b0 = KEY_NONE, b1 = KEY_NONE;
    const uint32_t* p = partial + (size_t)q * nparts * 2;
    for (int i = lane; i < nparts * 2; i += 32) top2_insert(b0, b1, p[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t o0 = __shfl_xor_sync(0xffffffffu, b0, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o);
        top2_insert(b0, b1, o0);
        top2_insert(b0, b1, o1);
    }
    if (lane == 0) {
        const int i0 = b0 == KEY_NONE ? -1 : (int)(b0 & ((1u << KEY_SHIFT) - 1)) + train_offset;
        const int i1 = b1 == KEY_NONE ? -1 : (int)(b1 & ((1u << KEY_SHIFT) - 1)) + train_offset;
        const int d0 = b0 == KEY_NONE ? -1 : (int)(b0 >> KEY_SHIFT), d1 = b1 == KEY_NONE ? -1 : (int)(b1 >> KEY_SHIFT);
        idx2[2 * q] = i0; idx2[2 * q + 1] = i1;
        dist2[2 * q] = d0; dist2[2 * q + 1] = d1;
        if (match) {
            // Frame.cc:1562  `(*it)[0].distance < (*it)[1].distance * 0.7` (float * double)
            const bool ok = i0 >= 0 && i1 >= 0 && (double)(float)d0 < (double)(float)d1 * 0.7;
            match[q] = ok ? i0 : -1;
        }
    }
}

// Merge G per-shard (idx2, dist2) tables [G][nq][2] with global indices: order (dist, idx).
__global__ void k_knn2_merge_tables(const int32_t* __restrict__ idxS, const int32_t* __restrict__ distS,
                                    size_t shardStride, int G, int nq, int32_t* __restrict__ idx2,
                                    int32_t* __restrict__ dist2, int32_t* __restrict__ match) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned long long b0 = ~0ull, b1 = ~0ull;
    for (int s = 0; s < G; s++)
        for (int k = 0; k < 2; k++) {
            const int i = idxS[s * shardStride + (size_t)q * 2 + k], d = distS[s * shardStride + (size_t)q * 2 + k];
            if (i < 0) continue;
            const unsigned long long key = ((unsigned long long)(unsigned)d << 32) | (unsigned)i;
            const unsigned long long hi = b0 > key ? b0 : key;
            b0 = b0 < key ? b0 : key;
            b1 = b1 < hi ? b1 : hi;
        }
    const int i0 = b0 == ~0ull ? -1 : (int)(b0 & 0xFFFFFFFFu), i1 = b1 == ~0ull ? -1 : (int)(b1 & 0xFFFFFFFFu);
    const int d0 = b0 == ~0ull ? -1 : (int)(b0 >> 32), d1 = b1 == ~0ull ? -1 : (int)(b1 >> 32);
    idx2[2 * q] = i0; idx2[2 * q + 1] = i1;
    dist2[2 * q] = d0; dist2[2 * q + 1] = d1;
    if (match) match[q] = (i0 >= 0 && i1 >= 0 && (double)(float)d0 < (double)(float)d1 * 0.7) ? i0 : -1;
}

// The same merge with the exchange fused in: shard s's packed table { idx2[nq][2], dist2[nq][2] } is read where GPU s
// wrote it, through a peer mapping over NVLink (symmetric memory / CUDA IPC), so no all-gather runs before the merge.
constexpr int kMaxPeers = 16;
struct PeerTabs { const int32_t* t[kMaxPeers]; };

__global__ void k_knn2_merge_peers(PeerTabs P, int G, int nq, int32_t* __restrict__ idx2, int32_t* __restrict__ dist2,
                                   int32_t* __restrict__ match) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned long long b0 = ~0ull, b1 = ~0ull;
#pragma unroll 1
    for (int s = 0; s < G; s++) {
        const int32_t* tab = P.t[s];
        // one 8-byte peer load per table and query
        const int2 ii = *reinterpret_cast<const int2*>(tab + (size_t)q * 2);
        const int2 dd = *reinterpret_cast<const int2*>(tab + 2 * (size_t)nq + (size_t)q * 2);
        const int iv[2] = {ii.x, ii.y}, dv[2] = {dd.x, dd.y};
        for (int k = 0; k < 2; k++) {
            if (iv[k] < 0) continue;
            const unsigned long long key = ((unsigned long long)(unsigned)dv[k] << 32) | (unsigned)iv[k];
            const unsigned long long hi = b0 > key ? b0 : key;
            b0 = b0 < key ? b0 : key;
            b1 = b1 < hi ? b1 : hi;
        }
    }
    const int i0 = b0 == ~0ull ? -1 : (int)(b0 & 0xFFFFFFFFu), i1 = b1 == ~0ull ? -1 : (int)(b1 & 0xFFFFFFFFu);
    const int d0 = b0 == ~0ull ? -1 : (int)(b0 >> 32), d1 = b1 == ~0ull ? -1 : (int)(b1 >> 32);
    idx2[2 * q] = i0; idx2[2 * q + 1] = i1;
    dist2[2 * q] = d0; dist2[2 * q + 1] = d1;
    if (match) match[q] = (i0 >= 0 && i1 >= 0 && (double)(float)d0 < (double)(float)d1 * 0.7) ? i0 : -1;
}

// Frame::ComputeStereoFishEyeMatches' brute force (Frame.cc:1545-1562) for a BATCH of stereo pairs resident in HBM:
// pair b matches rows [qBegin[b], qEnd[b]) of its left descriptor slab (the lapping-area keypoints: monoIndex .. n)
// against rows [tBegin[b], tEnd[b]) of the right one.  grid = (query blocks, pairs); a CTA keeps 256 queries in
// registers and walks the pair's whole train range (a frame holds a few thousand rows: no partial tables to merge).
// Indices are relative to the range starts, as cv::BFMatcher numbers the rows of the two sub-matrices.
__global__ void __launch_bounds__(KNN_THREADS)
k_knn2_batch(const uint32_t* __restrict__ descQ, const int* __restrict__ qBegin, const int* __restrict__ qEnd,
             const uint32_t* __restrict__ descT, const int* __restrict__ tBegin, const int* __restrict__ tEnd, int capacity,
             int32_t* __restrict__ idx2, int32_t* __restrict__ dist2, int32_t* __restrict__ match) {
    __shared__ __align__(16) uint32_t tile[KNN_TILE * 8];
    const size_t b = blockIdx.y, slab = b * (size_t)capacity;
    const int qb = max(qBegin[b], 0), nq = min(qEnd[b], capacity) - qb;
    const int tb = max(tBegin[b], 0), nt = min(tEnd[b], capacity) - tb;
    if ((int)(blockIdx.x * KNN_QB) >= nq) return;          // uniform per CTA
    const uint32_t* query = descQ + 8 * (slab + qb);
    const uint32_t* train = descT + 8 * (slab + tb);
    const int q0 = blockIdx.x * KNN_QB + threadIdx.x;
    uint32_t qv[KNN_QPT][8];
    uint32_t b0[KNN_QPT], b1[KNN_QPT];
#pragma unroll
    for (int r = 0; r < KNN_QPT; r++) {
        const int q = min(q0 + r * KNN_THREADS, nq - 1);
        const uint4* p = reinterpret_cast<const uint4*>(query + 8 * (size_t)q);
        *reinterpret_cast<uint4*>(qv[r]) = p[0];
        *reinterpret_cast<uint4*>(qv[r] + 4) = p[1];
        b0[r] = KEY_NONE;
        b1[r] = KEY_NONE;
    }
    for (int t0 = 0; t0 < nt; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, nt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += KNN_THREADS)
            reinterpret_cast<uint4*>(tile)[i] = reinterpret_cast<const uint4*>(train + 8 * (size_t)t0)[i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            uint32_t tv[8];
            *reinterpret_cast<uint4*>(tv) = reinterpret_cast<const uint4*>(tile)[2 * j];
            *reinterpret_cast<uint4*>(tv + 4) = reinterpret_cast<const uint4*>(tile)[2 * j + 1];
            const uint32_t jj = (uint32_t)(t0 + j);
#pragma unroll
            for (int r = 0; r < KNN_QPT; r++) {
                const uint32_t key = ((uint32_t)hamming256_csa(qv[r], tv) << KEY_SHIFT) | jj;
                top2_insert(b0[r], b1[r], key);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KNN_QPT; r++) {
        const int q = q0 + r * KNN_THREADS;
        if (q >= nq) continue;
        const uint32_t k0 = b0[r], k1 = b1[r];
        const int i0 = k0 == KEY_NONE ? -1 : (int)(k0 & ((1u << KEY_SHIFT) - 1)), i1 = k1 == KEY_NONE ? -1 : (int)(k1 & ((1u << KEY_SHIFT) - 1));
        const int d0 = k0 == KEY_NONE ? -1 : (int)(k0 >> KEY_SHIFT), d1 = k1 == KEY_NONE ? -1 : (int)(k1 >> KEY_SHIFT);
        const size_t o = slab + q;
        idx2[2 * o] = i0; idx2[2 * o + 1] = i1;
        dist2[2 * o] = d0; dist2[2 * o + 1] = d1;
        // Frame.cc:1562  `(*it)[0].distance < (*it)[1].distance * 0.7` (float * double)
        if (match) match[o] = (i0 >= 0 && i1 >= 0 && (double)(float)d0 < (double)(float)d1 * 0.7) ? i0 : -1;
    }
}

int mfail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define MCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return mfail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

int set_device(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return mfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", e);
    if (device < 0 || device >= ndev) return mfail(ORBFE_ERR_INVALID, "bad device ordinal");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return mfail(ORBFE_ERR_CUDA, "cudaSetDevice", e);
    return ORBFE_OK;
}

}  // namespace

// Train descriptors per CTA: small enough that the grid covers the 148 SMs several times (a shard of
// a map split over 8 GPUs is only 125 k descriptors), a multiple of the staging tile.
static int knn2_chunk(int nq, int nt) {
    const int qblocks = (nq + KNN_QB - 1) / KNN_QB;
    const int wantChunks = (148 * 8 + qblocks - 1) / qblocks;
    int chunk = (nt + wantChunks - 1) / std::max(wantChunks, 1);
    chunk = (chunk + KNN_TILE - 1) / KNN_TILE * KNN_TILE;
    return std::min(std::max(chunk, KNN_TILE), KNN_CHUNK_MAX);
}

// Enqueue kNN-2 on `st`: idx2/dist2 (and match when non-null) for nq queries against nt train rows.
int orbfe_knn2_enqueue(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, int train_offset,
                       int32_t* d_idx2, int32_t* d_dist2, int32_t* d_match, uint32_t* d_partial,
                       cudaStream_t st) {
    const int chunk = knn2_chunk(nq, nt);
    int nchunks = std::max(1, (nt + chunk - 1) / chunk);
    if (orbfe_knn2_umma_parts(nq, nt) > 0) {
        // large maps: the same partial tables from the tensor cores (knn_umma.cu), same merge
        nchunks = orbfe_knn2_umma_enqueue((const uint32_t*)d_query, nq, (const uint32_t*)d_train, nt, d_partial, st);
        if (nchunks < 0) return nchunks;
    } else if (nt > 0) {
        dim3 grid((nq + KNN_QB - 1) / KNN_QB, nchunks);
        k_knn2_partial<<<grid, KNN_THREADS, 0, st>>>((const uint32_t*)d_query, nq, (const uint32_t*)d_train, nt,
                                                     chunk, nchunks, d_partial);
    } else {
        cudaMemsetAsync(d_partial, 0xFF, sizeof(uint32_t) * 2 * (size_t)nq, st);
    }
    k_knn2_merge_keys<<<(nq + 7) / 8, 256, 0, st>>>(d_partial, nq, nchunks, train_offset, d_idx2, d_dist2, d_match);
    return 2;
}

size_t orbfe_knn2_partial_bytes(int nq, int nt) {
    const int chunk = knn2_chunk(nq, nt);
    const int nchunks = std::max(std::max(1, (nt + chunk - 1) / chunk), orbfe_knn2_umma_parts(nq, nt));
    return sizeof(uint32_t) * 2 * (size_t)nq * nchunks;
}

extern "C" {

int orbfe_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* out, int device) {
    int rc = set_device(device);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!a || !b || !out))) return mfail(ORBFE_ERR_INVALID, "bad arguments");
    if (n == 0) return ORBFE_OK;
    OrbfeStage S;
    const size_t ia = S.in(a, 32 * (size_t)n), ib = S.in(b, 32 * (size_t)n), io = S.out(out, 4 * (size_t)n);
    MCK(S.commit(device));
    MCK(S.upload());
    k_hamming_pairs<<<(n + 255) / 256, 256, 0, S.stream()>>>(S.ptr<uint32_t>(ia), S.ptr<uint32_t>(ib), n, S.ptr<int32_t>(io));
    MCK(cudaGetLastError());
    MCK(S.download());
    return ORBFE_OK;
}

int orbfe_knn2_device(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, int train_offset,
                      int32_t* d_idx2, int32_t* d_dist2, void* stream) {
    if (nq <= 0) return ORBFE_OK;
    if (nt < 0 || nt >= (1 << KEY_SHIFT) || !d_query || !d_idx2 || !d_dist2)
        return mfail(ORBFE_ERR_INVALID, "bad arguments (nt must be < 2^23 per call: shard larger maps)");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* partial = nullptr;
    {   // the scratch comes from the device's stream-ordered pool: keep freed blocks in the pool across synchronisations
        // (the default threshold of 0 hands them back to the driver at every sync, and every call pays a fresh allocation)
        static bool kept[64] = {};
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && dev < 64 && !kept[dev] && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = 1ull << 30, cur = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < keep)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            kept[dev] = true;
        }
    }
    MCK(cudaMallocAsync((void**)&partial, orbfe_knn2_partial_bytes(nq, nt), st));
    const int erc = orbfe_knn2_enqueue(d_query, nq, d_train, nt, train_offset, d_idx2, d_dist2, nullptr, partial, st);
    if (erc < 0) { cudaFreeAsync(partial, st); return erc; }
    MCK(cudaGetLastError());
    MCK(cudaFreeAsync(partial, st));
    return ORBFE_OK;
}

int orbfe_knn2_batch_device(const uint8_t* d_desc_q, const int* d_q_begin, const int* d_q_end, const uint8_t* d_desc_t,
                            const int* d_t_begin, const int* d_t_end, int B, int capacity, int32_t* d_idx2, int32_t* d_dist2,
                            int32_t* d_match, void* stream) {
    if (B <= 0) return ORBFE_OK;
    if (capacity <= 0 || capacity >= (1 << KEY_SHIFT) || !d_desc_q || !d_q_begin || !d_q_end || !d_desc_t || !d_t_begin || !d_t_end ||
        !d_idx2 || !d_dist2)
        return mfail(ORBFE_ERR_INVALID, "bad arguments");
    const int urc = orbfe_knn2_umma_batch_enqueue((const uint32_t*)d_desc_q, d_q_begin, d_q_end, (const uint32_t*)d_desc_t, d_t_begin, d_t_end, B,
                                                  capacity, d_idx2, d_dist2, d_match, (cudaStream_t)stream);
    if (urc < 0) return urc;
    if (urc > 0) {      // frames of >= 512 rows: the tensor-core kernel (knn_umma.cu)
        MCK(cudaGetLastError());
        return ORBFE_OK;
    }
    k_knn2_batch<<<dim3((capacity + KNN_QB - 1) / KNN_QB, B), KNN_THREADS, 0, (cudaStream_t)stream>>>(
        (const uint32_t*)d_desc_q, d_q_begin, d_q_end, (const uint32_t*)d_desc_t, d_t_begin, d_t_end, capacity, d_idx2, d_dist2,
        d_match);
    MCK(cudaGetLastError());
    return ORBFE_OK;
}

int orbfe_knn2(const uint8_t* query, int nq, const uint8_t* train, int nt, int train_offset, int32_t* idx2,
               int32_t* dist2, int32_t* match, int device) {
    int rc = set_device(device);
    if (rc) return rc;
    if (nq < 0 || nt < 0 || nt >= (1 << KEY_SHIFT)) return mfail(ORBFE_ERR_INVALID, "bad sizes (nt must be < 2^23 per call)");
    if (nq == 0) return ORBFE_OK;
    if (!query || (nt > 0 && !train) || !idx2 || !dist2) return mfail(ORBFE_ERR_INVALID, "null argument");
    OrbfeStage S;
    const size_t iq = S.in(query, 32 * (size_t)nq), it = S.in(train, 32 * (size_t)nt);
    const size_t wp = S.work(orbfe_knn2_partial_bytes(nq, nt));
    const size_t oi = S.out(idx2, 8 * (size_t)nq), od = S.out(dist2, 8 * (size_t)nq), om = S.out(match, 4 * (size_t)nq);
    MCK(S.commit(device));
    MCK(S.upload());
    rc = orbfe_knn2_enqueue(S.ptr<uint8_t>(iq), nq, S.ptr<uint8_t>(it), nt, train_offset, S.ptr<int32_t>(oi), S.ptr<int32_t>(od),
                            S.ptr<int32_t>(om), S.ptr<uint32_t>(wp), S.stream());
    if (rc < 0) return rc;
    MCK(cudaGetLastError());
    MCK(S.download());
    return ORBFE_OK;
}

int orbfe_knn2_merge_device(const int32_t* d_idx2_shards, const int32_t* d_dist2_shards, int G, int nq,
                            int32_t* d_idx2, int32_t* d_dist2, int32_t* d_match, void* stream) {
    if (nq <= 0 || G <= 0) return ORBFE_OK;
    k_knn2_merge_tables<<<(nq + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_idx2_shards, d_dist2_shards, 2 * (size_t)nq, G,
                                                                           nq, d_idx2, d_dist2, d_match);
    MCK(cudaGetLastError());
    return ORBFE_OK;
}

int orbfe_knn2_merge_packed_device(const int32_t* d_packed, int G, int nq, int32_t* d_idx2, int32_t* d_dist2,
                                   int32_t* d_match, void* stream) {
    if (nq <= 0 || G <= 0) return ORBFE_OK;
    // packed[s] = { idx2[nq][2], dist2[nq][2] } of shard s: one all-gather moves both tables
    k_knn2_merge_tables<<<(nq + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_packed, d_packed + 2 * (size_t)nq, 4 * (size_t)nq,
                                                                           G, nq, d_idx2, d_dist2, d_match);
    MCK(cudaGetLastError());
    return ORBFE_OK;
}

int orbfe_knn2_merge_peers_device(const int32_t* const* peer_tabs, int G, int nq, int32_t* d_idx2, int32_t* d_dist2,
                                  int32_t* d_match, void* stream) {
    if (nq <= 0 || G <= 0) return ORBFE_OK;
    if (!peer_tabs || !d_idx2 || !d_dist2) return mfail(ORBFE_ERR_INVALID, "null argument");
    if (G > kMaxPeers) return mfail(ORBFE_ERR_CAPACITY, "more than 16 shards: gather the tables and use orbfe_knn2_merge_packed_device");
    PeerTabs P;
    for (int s = 0; s < kMaxPeers; s++) P.t[s] = s < G ? peer_tabs[s] : nullptr;
    for (int s = 0; s < G; s++)
        if (!P.t[s]) return mfail(ORBFE_ERR_INVALID, "null peer table");
    k_knn2_merge_peers<<<(nq + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, G, nq, d_idx2, d_dist2, d_match);
    MCK(cudaGetLastError());
    return ORBFE_OK;
}

int orbfe_knn2_merge(const int32_t* idx2_shards, const int32_t* dist2_shards, int G, int nq, int32_t* idx2,
                     int32_t* dist2, int32_t* match, int device) {
    int rc = set_device(device);
    if (rc) return rc;
    if (nq <= 0 || G <= 0) return ORBFE_OK;
    const size_t tb = 8 * (size_t)nq * G;
    OrbfeStage S;
    const size_t si = S.in(idx2_shards, tb), sd = S.in(dist2_shards, tb);
    const size_t oi = S.out(idx2, 8 * (size_t)nq), od = S.out(dist2, 8 * (size_t)nq), om = S.out(match, 4 * (size_t)nq);
    MCK(S.commit(device));
    MCK(S.upload());
    rc = orbfe_knn2_merge_device(S.ptr<int32_t>(si), S.ptr<int32_t>(sd), G, nq, S.ptr<int32_t>(oi), S.ptr<int32_t>(od),
                                 S.ptr<int32_t>(om), S.stream());
    if (rc) return rc;
    MCK(S.download());
    return ORBFE_OK;
}

}  // extern "C"
