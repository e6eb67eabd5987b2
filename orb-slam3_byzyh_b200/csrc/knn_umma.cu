// knn_umma.cu -- the brute-force Hamming kNN-2 of match.cu on the 5th-generation tensor cores, for maps large enough to
// fill the machine (cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2), /root/reference/src/Frame.cc:1553, and BASELINE config 5:
// 2000 frame descriptors against 1 M map descriptors).
//
// Every descriptor bit becomes one signed byte, +1 for a set bit and -1 for a clear one, so the int8 dot product of two
// descriptors is S = 256 - 2 * hamming: a (queries x 256) . (256 x map points) contraction, tcgen05.mma kind::i8 with
// the accumulator in tensor memory.  A CTA owns 128 queries (operand A, expanded once) and a range of the map.  Eight
// producer warps expand map tiles of 256 points from their bits (32 B per point in HBM, 256 B in shared memory, written in
// the 8-row x 16-byte core-matrix layout the MMA reads, no swizzle), one thread issues 8 x (128 x 256 x 32) MMAs per tile
// into one of two 256-column accumulators, and eight warps (four lane quadrants x two column halves) read the accumulators
// back with tcgen05.ld (.pack::16b: the value fits 16 bits, which halves the register traffic of the read) and keep the
// best two keys per query.  key = distance << 23 | map index, the order of match.cu (ties -> lowest index), so the partial
// tables go through the same merge kernel and the result is bit-identical to the scalar kernel's.
//
// The roofline of this kernel is the accumulator read: 4 B of tensor memory per descriptor pair at 64 B / clk / SM
// (measured figure of the microarchitecture guide), ahead of the MMA itself (1084 clk per 128 x 256 tile) -- DESIGN.md.
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>

#include "orbfe_internal.h"

namespace {

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int KEY_SHIFT = 23;
constexpr int QT = 128;                 // queries per CTA = UMMA M
constexpr int MT = 256;                 // map points per tile = UMMA N
constexpr int KBYTES = 256;             // one signed byte per descriptor bit
constexpr int A_BYTES = QT * KBYTES, B_BYTES = MT * KBYTES;
constexpr int EPI_WARPS = 8, PROD_WARPS = 8, NWARPS = EPI_WARPS + 1 + PROD_WARPS;   // epilogue: 2 column halves x 4 lane quadrants
constexpr int SMEM_BYTES = A_BYTES + 2 * B_BYTES + 128 + QT * 8 + 1024;
// operand tiles are stored as 8-row x 16-byte core matrices (128 contiguous bytes), K chunks next to each other:
// core (row group g, K chunk c) at (g * 16 + c) * 128
constexpr uint32_t CORE = 128, LBO = CORE, SBO = 16 * CORE;
static_assert(PROD_WARPS * 32 == MT, "one producer thread per row of the map tile");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {   // bounded (2 s): a trap is better than a hung GPU
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = global_ns();
    while (!mbar_try_wait(bar, parity))
        if (global_ns() - t0 > 2000000000ull) __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// 64 accumulator columns as 32 registers: the low 16 bits of two adjacent columns per register (|S| <= 256)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void top2_insert_s(int& b0, int& b1, int k) {
    const int hi = max(b0, k);
    b0 = min(b0, k);
    b1 = min(b1, hi);
}

// the registers of the load are operands of the wait, so that nothing that reads them can move above it
__device__ __forceinline__ void tmem_wait_ld(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// 64 accumulator columns (first column index c0 inside the tile; register r = columns 2 r, 2 r + 1 as signed 16-bit
// halves) against the kept keys.  key' = -S * 2^22 + column: a column can only enter the best two if S > sThr =
// floor(-b1 / 2^22), so the common case is a tree of packed 16-bit max and one compare.  thr2 = sThr in both halves.
__device__ __forceinline__ uint32_t pack_thr(int sThr) {
    const uint32_t t = (uint32_t)max(min(sThr, 32767), -32768) & 0xFFFFu;
    return t | (t << 16);
}
__device__ __forceinline__ void epi_chunk(const uint32_t* v, int c0, int valid, int& b0, int& b1, uint32_t& thr2, bool live) {
    if (valid == MT) {
        uint32_t m[8];
#pragma unroll
        for (int g = 0; g < 8; g++) m[g] = __vmaxs2(__vimax3_s16x2(v[4 * g], v[4 * g + 1], v[4 * g + 2]), v[4 * g + 3]);
        const uint32_t mAll = __vimax3_s16x2(__vimax3_s16x2(m[0], m[1], m[2]), __vimax3_s16x2(m[3], m[4], m[5]), __vmaxs2(m[6], m[7]));
        if (__vmaxs2(mAll, thr2) != thr2) {
#pragma unroll
            for (int g = 0; g < 8; g++) {
                if (__vmaxs2(m[g], thr2) != thr2) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int sLo = (int)(short)(v[4 * g + i] & 0xFFFFu), sHi = (int)v[4 * g + i] >> 16;
                        top2_insert_s(b0, b1, sLo * -(1 << 22) + (c0 + 8 * g + 2 * i));
                        top2_insert_s(b0, b1, sHi * -(1 << 22) + (c0 + 8 * g + 2 * i + 1));
                    }
                    thr2 = pack_thr((-b1) >> 22);
                }
            }
        }
    } else if (live) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const int sLo = (int)(short)(v[j] & 0xFFFFu), sHi = (int)v[j] >> 16;
            if (c0 + 2 * j < valid) top2_insert_s(b0, b1, sLo * -(1 << 22) + (c0 + 2 * j));
            if (c0 + 2 * j + 1 < valid) top2_insert_s(b0, b1, sHi * -(1 << 22) + (c0 + 2 * j + 1));
        }
        thr2 = pack_thr((-b1) >> 22);
    }
}

// K-major operand descriptor, no swizzle: start address, leading (K direction) and stride (row group) byte offsets in
// units of 16 bytes, descriptor version 1 (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// rows [0, rows) of an operand tile from descriptor bits (8 words per row), +1 for a set bit, -1 for a clear one; rows
// >= valid are zero.  Any fixed assignment of bits to K positions serves (both operands use this one): bit 4 j + i of a
// word goes to byte j of output word i, so (w >> i) & 0x11111111 is at once the PRMT selector that picks 0x01 or 0xFF.
__device__ __forceinline__ void expand_row(uint8_t* dst, int p, uint4 w0, uint4 w1, bool ok) {
    uint8_t* rb = dst + (p >> 3) * SBO + (p & 7) * 16;
    if (ok) {
        const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t t = (w[i] >> k) & 0x11111111u;
                lo[k] = __byte_perm(0x000001FFu, 0u, t);
                hi[k] = __byte_perm(0x000001FFu, 0u, t >> 16);
            }
            *reinterpret_cast<uint4*>(rb + (2 * i) * LBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(rb + (2 * i + 1) * LBO) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) *reinterpret_cast<uint4*>(rb + i * LBO) = make_uint4(0u, 0u, 0u, 0u);
    }
}
__device__ __forceinline__ void expand_rows(uint8_t* dst, const uint32_t* __restrict__ src, int rows, int valid, int tid, int nthr) {
    for (int p = tid; p < rows; p += nthr) {
        uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;
        if (p < valid) { w0 = __ldg(reinterpret_cast<const uint4*>(src + 8 * (size_t)p)); w1 = __ldg(reinterpret_cast<const uint4*>(src + 8 * (size_t)p) + 1); }
        expand_row(dst, p, w0, w1, p < valid);
    }
}


// grid = (query tiles, map splits); partial[q][split][2] = the two smallest (distance << 23 | map index) keys
// Batch form (kBatch): blockIdx.y = stereo pair; the pair's query / train row ranges come from device arrays (as in
// k_knn2_batch of match.cu), one CTA walks the pair's whole train range, and the two column halves are merged inside the
// CTA, so idx2 / dist2 / match are written directly.
struct KnnBatchArgs {
    const int *qBegin, *qEnd, *tBegin, *tEnd;
    int capacity;
    int32_t *idx2, *dist2, *match;
};

template <bool kBatch>
__global__ void __launch_bounds__(32 * NWARPS, 1)
k_knn2_umma(const uint32_t* __restrict__ query, int nq, const uint32_t* __restrict__ train, int nt, int nsplit,
            uint32_t* __restrict__ partial, const KnnBatchArgs ba) {
    size_t slab = 0;
    if constexpr (kBatch) {
        const int b = blockIdx.y;
        slab = (size_t)b * ba.capacity;
        const int qb = max(ba.qBegin[b], 0), tb = max(ba.tBegin[b], 0);
        nq = min(ba.qEnd[b], ba.capacity) - qb;
        nt = max(min(ba.tEnd[b], ba.capacity) - tb, 0);
        if ((int)(blockIdx.x * QT) >= nq) return;          // uniform per CTA, before anything is allocated
        query += 8 * (slab + qb);
        train += 8 * (slab + tb);
        nsplit = 1;
    }
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + 2 * B_BYTES);
    uint64_t *bFull = bars, *bEmpty = bars + 2, *accFull = bars + 4, *accEmpty = bars + 6;
    uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bars + 8);
    uint32_t* halfKeys = reinterpret_cast<uint32_t*>(bars + 10);        // batch form: [QT][2] keys of column half 1
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // map tiles of this CTA
    const int tilesAll = (nt + MT - 1) / MT;
    const int split = kBatch ? 0 : (int)blockIdx.y;
    const int tile0 = (int)((long long)tilesAll * split / nsplit), tile1 = (int)((long long)tilesAll * (split + 1) / nsplit);
    const int ntiles = tile1 - tile0;
    const int q0 = blockIdx.x * QT;

    if (tid == 0) {
        mbar_init(&bFull[0], 32 * PROD_WARPS); mbar_init(&bFull[1], 32 * PROD_WARPS);
        mbar_init(&bEmpty[0], 1); mbar_init(&bEmpty[1], 1);
        mbar_init(&accFull[0], 1); mbar_init(&accFull[1], 1);
        mbar_init(&accEmpty[0], 32 * EPI_WARPS); mbar_init(&accEmpty[1], 32 * EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // the whole tensor memory: two 256-column accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemPtr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    expand_rows(sA, query + 8 * (size_t)q0, QT, min(QT, nq - q0), tid, 32 * NWARPS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemPtr;
    uint32_t mine0 = KEY_NONE, mine1 = KEY_NONE;     // batch form: keys of column half 0 of this thread's query

    if (warp < EPI_WARPS) {
        // ===== accumulators -> best two keys.  key' = -S * 2^22 + (column - tile base) = (distance - 128) * 2^23 + relative
        // index; after every tile the kept keys move by -256, so they stay relative to the base of the current tile =====
        // rows past the last query (their operand rows are zero) never take the update path
        const int quad = warp & 3, half = warp >> 2;      // TMEM lanes 32 quad .. 32 quad + 31, columns 128 half .. 128 half + 127
        const bool live = q0 + quad * 32 + lane < nq;
        int b0 = INT_MAX, b1 = INT_MAX;
        uint32_t thr2 = pack_thr(live ? INT_MIN : INT_MAX);
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1, ph = (t >> 1) & 1;
            mbar_wait(&accFull[s], ph);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * MT + half * (MT / 2));
            const int valid = min(MT, nt - (tile0 + t) * MT);
            const int cb = half * (MT / 2);
            uint32_t va[32], vb[32];
            tmem_ld32(taddr, va);
            tmem_ld32(taddr + 64, vb);
            tmem_wait_ld(va);
            tmem_wait_ld(vb);
            tc_fence_before();
            mbar_arrive(&accEmpty[s]);     // the accumulator is in registers: the next MMA may overwrite it
            epi_chunk(va, cb, valid, b0, b1, thr2, live);
            epi_chunk(vb, cb + 64, valid, b0, b1, thr2, live);
            if (b0 != INT_MAX) b0 -= MT;
            if (b1 != INT_MAX) { b1 -= MT; thr2 = pack_thr((-b1) >> 22); }   // (dead rows keep b1 == INT_MAX)
        }
        const int q = q0 + quad * 32 + lane;
        if (q < nq) {
            const int baseEnd = tile1 * MT;
            uint32_t o[2];
            const int b[2] = {b0, b1};
#pragma unroll
            for (int i = 0; i < 2; i++) {
                if (b[i] == INT_MAX) { o[i] = KEY_NONE; continue; }
                const int D = (b[i] + (1 << 22)) >> 23, rel = b[i] - D * (1 << 23);
                o[i] = ((uint32_t)(D + 128) << KEY_SHIFT) | (uint32_t)(baseEnd + rel);
            }
            if constexpr (kBatch) {
                if (half == 1) { halfKeys[2 * (quad * 32 + lane)] = o[0]; halfKeys[2 * (quad * 32 + lane) + 1] = o[1]; }
                else { mine0 = o[0]; mine1 = o[1]; }
            } else {
                partial[((size_t)q * nsplit * 2 + blockIdx.y * 2 + half) * 2] = o[0];
                partial[((size_t)q * nsplit * 2 + blockIdx.y * 2 + half) * 2 + 1] = o[1];
            }
        }
    } else if (warp == EPI_WARPS) {
        // ===== one thread issues the MMAs: 8 x (128 x 256 x 32) per tile =====
        if (lane == 0) {
            // instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, N = 256, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
            const uint64_t dA = umma_desc(smem_u32(sA), LBO, SBO);
            for (int t = 0; t < ntiles; t++) {
                const int s = t & 1, ph = (t >> 1) & 1;
                mbar_wait(&bFull[s], ph);
                mbar_wait(&accEmpty[s], ph ^ 1);
                tc_fence_after();
                const uint64_t dB = umma_desc(smem_u32(sB + s * B_BYTES), LBO, SBO);
#pragma unroll
                for (int k = 0; k < KBYTES / 32; k++)   // 32 K bytes = two 16-byte chunks per instruction
                    umma_i8(tmem + (uint32_t)(s * MT), dA + (uint64_t)((k * 2 * CORE) >> 4), dB + (uint64_t)((k * 2 * CORE) >> 4), idesc, k > 0);
                umma_commit(&bEmpty[s]);
                umma_commit(&accFull[s]);
            }
        }
    } else {
        // ===== producers: map tile bits -> signed bytes in the operand layout =====
        const int ptid = tid - 32 * (EPI_WARPS + 1);     // = row of the tile (PROD_WARPS * 32 == MT)
        uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;
        {
            const int m = tile0 * MT + ptid;
            if (ntiles > 0 && m < nt) { w0 = __ldg(reinterpret_cast<const uint4*>(train + 8 * (size_t)m)); w1 = __ldg(reinterpret_cast<const uint4*>(train + 8 * (size_t)m) + 1); }
        }
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1, ph = (t >> 1) & 1;
            const uint4 c0 = w0, c1 = w1;
            const bool ok = (tile0 + t) * MT + ptid < nt;
            {   // the bits of the next tile travel while this one is expanded
                const int m = (tile0 + t + 1) * MT + ptid;
                if (t + 1 < ntiles && m < nt) { w0 = __ldg(reinterpret_cast<const uint4*>(train + 8 * (size_t)m)); w1 = __ldg(reinterpret_cast<const uint4*>(train + 8 * (size_t)m) + 1); }
            }
            mbar_wait(&bEmpty[s], ph ^ 1);
            expand_row(sB + s * B_BYTES, ptid, c0, c1, ok);
            fence_proxy_async();
            mbar_arrive(&bFull[s]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kBatch) {
        const int q = q0 + warp * 32 + lane;
        if (warp < 4 && q < nq) {
            uint32_t k0 = mine0, k1 = mine1;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const uint32_t k = halfKeys[2 * (warp * 32 + lane) + i], hi = max(k0, k);
                k0 = min(k0, k);
                k1 = min(k1, hi);
            }
            const int i0 = k0 == KEY_NONE ? -1 : (int)(k0 & ((1u << KEY_SHIFT) - 1)), i1 = k1 == KEY_NONE ? -1 : (int)(k1 & ((1u << KEY_SHIFT) - 1));
            const int d0 = k0 == KEY_NONE ? -1 : (int)(k0 >> KEY_SHIFT), d1 = k1 == KEY_NONE ? -1 : (int)(k1 >> KEY_SHIFT);
            const size_t o = slab + q;
            ba.idx2[2 * o] = i0; ba.idx2[2 * o + 1] = i1;
            ba.dist2[2 * o] = d0; ba.dist2[2 * o + 1] = d1;
            // Frame.cc:1562  `(*it)[0].distance < (*it)[1].distance * 0.7` (float * double)
            if (ba.match) ba.match[o] = (i0 >= 0 && i1 >= 0 && (double)(float)d0 < (double)(float)d1 * 0.7) ? i0 : -1;
        }
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}



// map splits of a launch: about one CTA per SM over (query tiles x splits); a split spans < 2^22 map points (the kept
// keys are relative to the current tile) and at least one tile
int umma_splits(int nq, int nt) {
    const int qtiles = (nq + QT - 1) / QT, tiles = (nt + MT - 1) / MT;
    int ns = 148 / qtiles;
    const int need = (int)(((long long)nt + (1 << 22) - MT - 1) / ((1 << 22) - MT));
    if (ns < need) ns = need;
    if (ns < 1) ns = 1;
    if (ns > tiles) ns = tiles;
    return ns;
}

}  // namespace

// Number of (key0, key1) pairs per query the tensor-core kernel writes for this problem size, 0 when the size is left to
// the scalar kernel of match.cu (small problems: a CTA here holds the whole tensor memory and 160 KB of shared memory).
// ORBFE_KNN_SCALAR=1 forces the scalar kernel (comparisons).
int orbfe_knn2_umma_parts(int nq, int nt) {
    static const bool off = getenv("ORBFE_KNN_SCALAR") && atoi(getenv("ORBFE_KNN_SCALAR")) != 0;
    if (off || nq <= 0 || nt < 512 || (long long)nq * nt < (1LL << 18)) return 0;
    return 2 * umma_splits(nq, nt);
}

// partial[q][part][2]; returns the number of parts per query, or a negative ORBFE error code.
int orbfe_knn2_umma_enqueue(const uint32_t* d_query, int nq, const uint32_t* d_train, int nt, uint32_t* d_partial, cudaStream_t st) {
    static bool attr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return orbfe_fail(ORBFE_ERR_CUDA, "cudaGetDevice", cudaGetLastError());
    if (dev < 64 && !attr[dev]) {
        if (cudaFuncSetAttribute(k_knn2_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_knn2_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
            return orbfe_fail(ORBFE_ERR_CUDA, "cudaFuncSetAttribute(k_knn2_umma)", cudaGetLastError());
        attr[dev] = true;
    }
    if (d_partial == nullptr) return 0;   // attribute set-up only (the batch entry point)
    const int ns = umma_splits(nq, nt);
    k_knn2_umma<false><<<dim3((nq + QT - 1) / QT, ns), 32 * NWARPS, SMEM_BYTES, st>>>(d_query, nq, d_train, nt, ns, d_partial, KnnBatchArgs{});
    return 2 * ns;
}

// The batched stereo form (orbfe_knn2_batch_device): 1 when the tensor-core kernel took the call, 0 when the sizes are left
// to the scalar kernel (frames with fewer than 512 descriptor rows), negative on error.
int orbfe_knn2_umma_batch_enqueue(const uint32_t* d_desc_q, const int* d_q_begin, const int* d_q_end, const uint32_t* d_desc_t,
                                  const int* d_t_begin, const int* d_t_end, int B, int capacity, int32_t* d_idx2, int32_t* d_dist2,
                                  int32_t* d_match, cudaStream_t st) {
    if (capacity < 512 || orbfe_knn2_umma_parts(capacity, capacity) == 0) return 0;
    const int rc = orbfe_knn2_umma_enqueue(nullptr, 0, nullptr, 0, nullptr, st);
    if (rc < 0) return rc;
    const KnnBatchArgs ba = {d_q_begin, d_q_end, d_t_begin, d_t_end, capacity, d_idx2, d_dist2, d_match};
    k_knn2_umma<true><<<dim3((capacity + QT - 1) / QT, B), 32 * NWARPS, SMEM_BYTES, st>>>(d_desc_q, 0, d_desc_t, 0, 1, nullptr, ba);
    return 1;
}
