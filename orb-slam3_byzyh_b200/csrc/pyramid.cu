// pyramid.cu -- ORBextractor::ComputePyramid (/root/reference/src/ORBextractor.cc:1687-1740).
//
// Level 0 is copyMakeBorder(image, 19 px, BORDER_REFLECT_101) (:1734-1736); level l > 0 is
// cv::resize(level l-1, INTER_LINEAR) into the centre of a (w+38) x (h+38) buffer followed by
// an in-place reflect-101 border (:1702-1716).  OpenCV is not vendored by the reference; the
// fixed-point bilinear arithmetic restated here (11-bit taps computed on the host in double,
// horizontal pass in int32, vertical pass (((b0*(H0>>4))>>16)+((b1*(H1>>4))>>16)+2)>>2) is the
// cv2 4.13 CV_8U path (oracle/cvprims.cpp, pinned against cv2 by tests/test_oracle_cvprims.py).
//
// Design: resize and border are ONE pass.  Every thread owns one aligned 4-byte word of the
// padded destination row; border pixels are produced by reflecting the destination coordinate
// into the ROI and evaluating the same bilinear tap there, so no second pass (and no
// synchronisation) is needed and every store is a coalesced 32-bit word.
#include <algorithm>

#include <cstdlib>

#include "orbfe_internal.h"
#include "remap_core.h"
#include "tma.h"

namespace {

__device__ __forceinline__ int reflect101_clamped(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * (len - 1) - p;
    return max(0, min(p, len - 1));
}

// Level 0 = copyMakeBorder(image, 19, REFLECT_101).  A thread owns 16 bytes of one padded row.  The 16-byte groups that
// lie entirely inside the image are one aligned 16-byte load + store (k_level0: interior groups only, so no warp ever
// takes the gather path); the two or three groups at either end of a row, which mix reflected border pixels, padding
// and image pixels, are gathered byte by byte by a second, small launch (k_level0_border).  With both in one kernel
// every warp contained a border group and paid the ~200-instruction gather (ncu: 0.46 M warp-instructions per frame
// for a 0.36 MB copy).
__device__ __forceinline__ uint4 level0_gather(const uint8_t* __restrict__ src, int x0, int w) {
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k] = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            v[k] |= (uint32_t)__ldg(src + reflect101_clamped(x0 + 4 * k + i, w)) << (8 * i);
    }
    return make_uint4(v[0], v[1], v[2], v[3]);
}

// gFirst .. gLast: the groups with 0 <= x0 and x0 + 15 < w
__global__ void __launch_bounds__(256)
k_level0(const uint8_t* __restrict__ img, size_t step, size_t frameStride, uint8_t* __restrict__ pyr,
         unsigned long long pyrStride, int w, int h, int pitch, int gFirst, int gLast) {
    const int gx = gFirst + blockIdx.x * 64 + (threadIdx.x & 63);  // 16-byte group inside the padded row
    const int py = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (gx > gLast || py >= h + 2 * ORBFE_YOFF) return;
    const uint8_t* src = img + (size_t)blockIdx.z * frameStride + (size_t)reflect101_clamped(py - ORBFE_YOFF, h) * step;
    const int x0 = 16 * gx - ORBFE_XOFF;
    uint4 out;
    if ((reinterpret_cast<size_t>(src + x0) & 15) == 0) out = __ldg(reinterpret_cast<const uint4*>(src + x0));
    else out = level0_gather(src, x0, w);                          // caller's rows are not 16-byte aligned
    *reinterpret_cast<uint4*>(pyr + (size_t)blockIdx.z * pyrStride + (size_t)py * pitch + 16 * gx) = out;
}

// The groups left of gFirst and right of gLast of every padded row: thread = (row, border group).
__global__ void __launch_bounds__(256)
k_level0_border(const uint8_t* __restrict__ img, size_t step, size_t frameStride, uint8_t* __restrict__ pyr,
                unsigned long long pyrStride, int w, int h, int pitch, int gFirst, int gLast) {
    const int groups = pitch >> 4, nb = gFirst + (groups - 1 - gLast);
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int py = t / nb, b = t - py * nb;
    if (py >= h + 2 * ORBFE_YOFF) return;
    const int gx = b < gFirst ? b : gLast + 1 + (b - gFirst);
    const uint8_t* src = img + (size_t)blockIdx.z * frameStride + (size_t)reflect101_clamped(py - ORBFE_YOFF, h) * step;
    *reinterpret_cast<uint4*>(pyr + (size_t)blockIdx.z * pyrStride + (size_t)py * pitch + 16 * gx) =
        level0_gather(src, 16 * gx - ORBFE_XOFF, w);
}

// Level 0 with the stereo rectification of System::TrackStereo fused in (src/System.cc:286-293): every padded pixel is
// cv::remap(INTER_LINEAR, CV_32FC1 maps, constant 0 border) evaluated at the reflect-101 image of its coordinate, so the
// rectified image never exists outside the pyramid.  Arithmetic: remap_core.h (bit-exact with cv2).
__global__ void __launch_bounds__(256)
k_level0_rect(const uint8_t* __restrict__ img, size_t step, size_t frameStride, int srows, int scols,
              const float* __restrict__ mapx, const float* __restrict__ mapy, uint8_t* __restrict__ pyr,
              unsigned long long pyrStride, int w, int h, int pitch) {
    const int wx = blockIdx.x * 64 + (threadIdx.x & 63);          // 4-byte word inside the padded row
    const int py = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (wx >= (pitch >> 2) || py >= h + 2 * ORBFE_YOFF) return;
    const uint8_t* src = img + (size_t)blockIdx.z * frameStride;
    const int y = reflect101_clamped(py - ORBFE_YOFF, h);
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = reflect101_clamped(4 * wx - ORBFE_XOFF + i, w);
        out |= orbfe_remap_sample(src, step, srows, scols, mapx[(size_t)y * w + x], mapy[(size_t)y * w + x]) << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(pyr + (size_t)blockIdx.z * pyrStride + (size_t)py * pitch + 4 * wx) = out;
}

// mode 0: bilinear taps; 1: exact 2x2 area average; 2: identity
template <int MODE>
__global__ void __launch_bounds__(256)
k_resize(uint8_t* pyr, unsigned long long pyrStride, const OrbfeTap* __restrict__ xtab,
         const OrbfeTap* __restrict__ ytab, unsigned srcOff, int srcPitch, unsigned dstOff, int w,
         int h, int pitch) {
    const int words = pitch >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int H = h + 2 * ORBFE_YOFF;
    if (idx >= words * H) return;
    const int py = idx / words, wc = idx - py * words;
    uint8_t* base = pyr + (size_t)blockIdx.y * pyrStride;
    const uint8_t* sroi = base + srcOff + (size_t)ORBFE_YOFF * srcPitch + ORBFE_XOFF;
    const int ry = reflect101_clamped(py - ORBFE_YOFF, h);
    const int x0 = 4 * wc - ORBFE_XOFF;
    uint32_t out = 0;
    if (MODE == 0) {
        const OrbfeTap ty = ytab[ry];
        const uint8_t* r0 = sroi + (size_t)ty.s * srcPitch;
        const uint8_t* r1 = sroi + (size_t)ty.s1 * srcPitch;
        const int b0 = ty.a0, b1 = ty.a1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const OrbfeTap tx = xtab[reflect101_clamped(x0 + i, w)];
            const int h0 = (int)r0[tx.s] * tx.a0 + (int)r0[tx.s1] * tx.a1;
            const int h1 = (int)r1[tx.s] * tx.a0 + (int)r1[tx.s1] * tx.a1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 255) << (8 * i);
        }
    } else if (MODE == 1) {
        const uint8_t* r0 = sroi + (size_t)(2 * ry) * srcPitch;
        const uint8_t* r1 = r0 + srcPitch;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int rx = 2 * reflect101_clamped(x0 + i, w);
            const int v = ((int)r0[rx] + (int)r0[rx + 1] + (int)r1[rx] + (int)r1[rx + 1] + 2) >> 2;
            out |= (uint32_t)v << (8 * i);
        }
    } else {
        const uint8_t* r0 = sroi + (size_t)ry * srcPitch;
#pragma unroll
        for (int i = 0; i < 4; i++) out |= (uint32_t)r0[reflect101_clamped(x0 + i, w)] << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(base + dstOff + (size_t)py * pitch + 4 * wc) = out;
}

// Fast bilinear path (the four source taps of four adjacent destination pixels span at most 8
// source bytes, i.e. scale <= 2; the host checks this when it builds the tap tables).  A thread
// owns EIGHT adjacent PADDED destination columns (two aligned words) and walks down RS_ROWS
// destination rows: the x-taps live in registers, every source row segment is fetched as three
// aligned 32-bit words per word of output, one PRMT per pixel pulls its two tap bytes out of that
// 8-byte window and one IDP2A applies the two 11-bit weights.  The horizontal pass of a source row
// is kept (already >> 4) and reused when the next destination row needs the same source row (at
// scale 1.2 that is 5 rows out of 6); the y-taps of the strip are fetched once and handed out by
// shuffle; the source row the NEXT destination row will need is requested before the current row
// is computed.  Border columns/rows are the same computation on the reflected destination
// coordinate, so resize + copyMakeBorder stay one pass with coalesced 64-bit stores.
constexpr int RS_ROWS = 16, RS_WARPS = 4, RS_G = 1;   // rows per warp strip, warps per CTA, words per thread
// (RS_G = 2 halves the tap setup per pixel but needs 77 registers: 24 resident warps per SM left the row loads exposed;
//  one word per thread runs in 48 registers, 40 warps per SM: 1.77 -> 1.60 ms per 1024 frames)

__global__ void __launch_bounds__(32 * RS_WARPS)
k_resize_fast(uint8_t* pyr, unsigned long long pyrStride, const OrbfeTap* __restrict__ xtab,
              const OrbfeTap* __restrict__ ytab, unsigned srcOff, int srcPitch, unsigned dstOff, int w, int h,
              int pitch) {
    const int words = pitch >> 2, pairs = words / RS_G;     // pitch is a multiple of 16: words % 4 == 0
    const int lane = threadIdx.x & 31;
    const bool active = blockIdx.x * 32 + lane < pairs;     // no early exit: the warp shuffles below
    const int wc0 = RS_G * min(blockIdx.x * 32 + lane, pairs - 1);
    const int H = h + 2 * ORBFE_YOFF;
    const int py0 = (blockIdx.y * RS_WARPS + (threadIdx.x >> 5)) * RS_ROWS;
    if (py0 >= H) return;
    uint8_t* base = pyr + (size_t)blockIdx.z * pyrStride;
    uint32_t wgt[RS_G][4], sel[RS_G][4];
    int wi0[RS_G], sh[RS_G];
#pragma unroll
    for (int g = 0; g < RS_G; g++) {
        OrbfeTap tp[4];
        int lo = 1 << 30;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            tp[i] = xtab[reflect101_clamped(4 * (wc0 + g) - ORBFE_XOFF + i, w)];
            lo = min(lo, (int)tp[i].s);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            wgt[g][i] = (uint32_t)(uint16_t)tp[i].a0 | ((uint32_t)(uint16_t)tp[i].a1 << 16);
            sel[g][i] = (uint32_t)(tp[i].s - lo) | ((uint32_t)(tp[i].s1 - lo) << 4);   // byte0 = src[s], byte1 = src[s1]
        }
        const int col = ORBFE_XOFF + lo;
        wi0[g] = col >> 2;       // the three words wi0..wi0+2 stay inside the source row (19-px border + padding)
        sh[g] = 8 * (col & 3);
    }
    const int srcWords = srcPitch >> 2;
    const uint32_t* __restrict__ sroi = reinterpret_cast<const uint32_t*>(base + srcOff + (size_t)ORBFE_YOFF * srcPitch);

    struct Raw { uint32_t w[RS_G][3]; };
    auto fetch = [&](int r) {
        Raw x;
        const uint32_t* rw = sroi + (size_t)r * srcWords;
#pragma unroll
        for (int g = 0; g < RS_G; g++) { x.w[g][0] = rw[wi0[g]]; x.w[g][1] = rw[wi0[g] + 1]; x.w[g][2] = rw[wi0[g] + 2]; }
        return x;
    };
    // horizontal pass of one fetched source row, already >> 4
    auto hpass = [&](const Raw& x, uint32_t (*Hs)[4]) {
#pragma unroll
        for (int g = 0; g < RS_G; g++) {
            const uint32_t lo8 = __funnelshift_r(x.w[g][0], x.w[g][1], sh[g]), hi8 = __funnelshift_r(x.w[g][1], x.w[g][2], sh[g]);
#pragma unroll
            for (int i = 0; i < 4; i++) Hs[g][i] = __dp2a_lo(wgt[g][i], __byte_perm(lo8, hi8, sel[g][i]), 0u) >> 4;
        }
    };

    OrbfeTap myTap = {0, 0, 0, 0};
    if (lane < RS_ROWS) myTap = ytab[reflect101_clamped(min(py0 + lane, H - 1) - ORBFE_YOFF, h)];
    const unsigned tapS = (uint32_t)(uint16_t)myTap.s | ((uint32_t)(uint16_t)myTap.s1 << 16);
    const unsigned tapA = (uint32_t)(uint16_t)myTap.a0 | ((uint32_t)(uint16_t)myTap.a1 << 16);

    int c0 = -1, c1 = -1;
    uint32_t C0[RS_G][4], C1[RS_G][4];   // cached horizontal passes of source rows c0, c1
    const int n = min(RS_ROWS, H - py0);
    uint32_t* dst = reinterpret_cast<uint32_t*>(base + dstOff + (size_t)py0 * pitch) + wc0;
    unsigned rs = __shfl_sync(0xffffffffu, tapS, 0);
    int pr = (int)(rs >> 16);            // prefetched source row (the one the first row needs as r1)
    Raw P = fetch(pr);
    for (int i = 0; i < n; i++, dst += words) {
        const int r0 = (int)(rs & 0xFFFFu), r1 = (int)(rs >> 16);
        const unsigned wa = __shfl_sync(0xffffffffu, tapA, i);
        const Raw got = P;
        const int gotRow = pr;
        if (i + 1 < n) {   // request the source row the next destination row will newly need
            rs = __shfl_sync(0xffffffffu, tapS, i + 1);
            pr = (int)(rs >> 16);
            P = fetch(pr);
        }
        // all threads of a warp share the row, hence r0/r1/c0/c1: the branches below are warp-uniform
        if (r0 != c0) {
            if (r0 == c1) {
#pragma unroll
                for (int g = 0; g < RS_G; g++)
#pragma unroll
                    for (int k = 0; k < 4; k++) C0[g][k] = C1[g][k];
            } else if (r0 == gotRow) {
                hpass(got, C0);
            } else {
                hpass(fetch(r0), C0);
            }
            c0 = r0;
        }
        if (r1 != c1) {
            if (r1 == c0) {
#pragma unroll
                for (int g = 0; g < RS_G; g++)
#pragma unroll
                    for (int k = 0; k < 4; k++) C1[g][k] = C0[g][k];
            } else if (r1 == gotRow) {
                hpass(got, C1);
            } else {
                hpass(fetch(r1), C1);
            }
            c1 = r1;
        }
        // ((b*(H>>4))>>16) == umulhi(b << 16, H >> 4) for the non-negative 11-bit weights
        const uint32_t b0 = wa << 16, b1 = wa & 0xFFFF0000u;
        uint32_t o[RS_G];
#pragma unroll
        for (int g = 0; g < RS_G; g++) {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = (__umulhi(b0, C0[g][k]) + __umulhi(b1, C1[g][k]) + 2u) >> 2;
            o[g] = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
        }
        if (active) *dst = o[0];
    }
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}

// Tiled bilinear resize: a CTA produces ORBFE_RZ_DW x ORBFE_RZ_DH pixels of the PADDED destination level (border
// columns / rows are the same computation on the reflected coordinate, as above).  The bounding box of the block's
// source taps arrives as ONE TMA tensor copy of the previous level (box origin 16-byte aligned, host-computed per
// block); the horizontal pass runs once per needed source row into shared memory (32-bit, low 4 bits cleared = the
// reference's H >> 4 kept in place), the vertical pass reads its two rows from there.  No row caching logic, no
// per-row address arithmetic and no global loads in the loops: 31 -> ~11 instructions per pixel.

template <int RZ_WARPS>
__global__ void __launch_bounds__(32 * RZ_WARPS)
k_resize_tile(uint8_t* __restrict__ pyr, unsigned long long pyrStride, const __grid_constant__ CUtensorMap srcMap,
              const OrbfeTap* __restrict__ xtab, const OrbfeTap* __restrict__ ytab, const OrbfeTap* __restrict__ xblk,
              const OrbfeTap* __restrict__ yblk, unsigned dstOff, int w, int h, int pitch, int boxW, int boxH, int tilesPerCta) {
    extern __shared__ __align__(128) uint8_t rzs[];
    const int srcBytes = (boxW * boxH + 127) & ~127;
    uint32_t* Hb = reinterpret_cast<uint32_t*>(rzs + 2 * srcBytes);          // [boxH][ORBFE_RZ_DW]; two TMA buffers before it
    // per destination row of the block: byte offsets of its two source rows inside Hb and the two weights << 12
    __shared__ __align__(16) uint4 ytile[ORBFE_RZ_DH];
    __shared__ __align__(8) uint64_t bar[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int words = pitch >> 2;
    const int Hp = h + 2 * ORBFE_YOFF;
    const int nyb = (Hp + ORBFE_RZ_DH - 1) / ORBFE_RZ_DH;
    const int yb0 = blockIdx.y * tilesPerCta, yb1 = min(yb0 + tilesPerCta, nyb);
    const int cLo = xblk[blockIdx.x].s;
    // A CTA walks `tilesPerCta` row blocks of one column block: the x taps are set up once, and the source box of the
    // next block is requested (second buffer) before the current one is processed, so nobody spins on the copy.
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init_fence();
        mbar_expect_tx(&bar[0], (uint32_t)(boxW * boxH));
        tma_tile_g2s(rzs, &srcMap, cLo, ORBFE_YOFF + yblk[yb0].s, blockIdx.z, &bar[0]);
    }
    // this thread's four destination columns: x taps in registers
    const int wc = min(blockIdx.x * 32 + lane, words - 1);   // lanes past the row end redo its last word
    uint32_t wgt[4], sel[4];
    int wi0, sh;
    {
        OrbfeTap tp[4];
        int lo = 1 << 30;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            tp[i] = xtab[reflect101_clamped(4 * wc - ORBFE_XOFF + i, w)];
            lo = min(lo, (int)tp[i].s);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            wgt[i] = (uint32_t)(uint16_t)tp[i].a0 | ((uint32_t)(uint16_t)tp[i].a1 << 16);
            sel[i] = (uint32_t)(tp[i].s - lo) | ((uint32_t)(tp[i].s1 - lo) << 4);   // byte0 = src[s], byte1 = src[s1]
        }
        const int col = ORBFE_XOFF + lo - cLo;   // column inside the staged box
        wi0 = col >> 2;
        sh = 8 * (col & 3);
    }
    const int bw4 = boxW >> 2;
    uint8_t* dbase = pyr + (size_t)blockIdx.z * pyrStride + dstOff;
    __syncthreads();                 // the barriers are initialised for everybody
    for (int yb = yb0; yb < yb1; yb++) {
        const int k = (yb - yb0) & 1;
        const OrbfeTap ybk = yblk[yb];
        const int rLo = ybk.s, nsrc = ybk.s1 - ybk.s + 1;
        const int py0 = yb * ORBFE_RZ_DH, n = min(ORBFE_RZ_DH, Hp - py0);
        if (threadIdx.x == 0 && yb + 1 < yb1) {
            // buffer k^1 was last read by the horizontal pass of block yb - 1 (two __syncthreads ago)
            fence_proxy_async();
            mbar_expect_tx(&bar[k ^ 1], (uint32_t)(boxW * boxH));
            tma_tile_g2s(rzs + (k ^ 1) * srcBytes, &srcMap, cLo, ORBFE_YOFF + yblk[yb + 1].s, blockIdx.z, &bar[k ^ 1]);
        }
        if (threadIdx.x < n) {
            const OrbfeTap t = ytab[reflect101_clamped(py0 + threadIdx.x - ORBFE_YOFF, h)];
            ytile[threadIdx.x] = make_uint4((uint32_t)(t.s - rLo) * (ORBFE_RZ_DW * 4), (uint32_t)(t.s1 - rLo) * (ORBFE_RZ_DW * 4),
                                            (uint32_t)(uint16_t)t.a0 << 12, (uint32_t)(uint16_t)t.a1 << 12);
        }
        mbar_wait(&bar[k], ((yb - yb0) >> 1) & 1);
        // horizontal pass of every staged source row
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(rzs + k * srcBytes) + wi0;
#pragma unroll 4
        for (int r = wid; r < nsrc; r += RZ_WARPS) {
            const uint32_t* rw = sw + r * bw4;
            const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
            const uint32_t lo8 = __funnelshift_r(w0, w1, sh), hi8 = __funnelshift_r(w1, w2, sh);
            uint4 hv;
            hv.x = __dp2a_lo(wgt[0], __byte_perm(lo8, hi8, sel[0]), 0u) & ~15u;
            hv.y = __dp2a_lo(wgt[1], __byte_perm(lo8, hi8, sel[1]), 0u) & ~15u;
            hv.z = __dp2a_lo(wgt[2], __byte_perm(lo8, hi8, sel[2]), 0u) & ~15u;
            hv.w = __dp2a_lo(wgt[3], __byte_perm(lo8, hi8, sel[3]), 0u) & ~15u;
            *reinterpret_cast<uint4*>(Hb + r * ORBFE_RZ_DW + 4 * lane) = hv;
        }
        __syncthreads();             // Hb and ytile are complete
        // vertical pass: ((b * (H >> 4)) >> 16) == umulhi(b << 12, H & ~15) for the non-negative 11-bit weights.  The two
        // products and the rounding constant meet in one three-input add; pixel pairs are packed before the >> 2 (every
        // sum is < 1024, so the byte wanted from each 16-bit half survives the shared shift)
        {
            uint8_t* drow = dbase + (size_t)(py0 + wid) * pitch + 4 * wc;
            const uint32_t hl = (uint32_t)__cvta_generic_to_shared(Hb + 4 * lane);
            const uint32_t yt = (uint32_t)__cvta_generic_to_shared(ytile);
#pragma unroll 4
            for (int i = wid; i < n; i += RZ_WARPS, drow += (size_t)RZ_WARPS * pitch) {
                const uint4 tq = lds128(yt + 16 * i);
                const uint4 h0 = lds128(hl + tq.x), h1 = lds128(hl + tq.y);
                // pairs first ((x1 << 16) + x0 is one LEA), then both products and the two rounding constants in one add
                const uint32_t a01 = __umulhi(tq.z, h0.x) + (__umulhi(tq.z, h0.y) << 16);
                const uint32_t b01 = __umulhi(tq.w, h1.x) + (__umulhi(tq.w, h1.y) << 16);
                const uint32_t a23 = __umulhi(tq.z, h0.z) + (__umulhi(tq.z, h0.w) << 16);
                const uint32_t b23 = __umulhi(tq.w, h1.z) + (__umulhi(tq.w, h1.w) << 16);
                const uint32_t lo = (a01 + b01 + 0x00020002u) >> 2, hi = (a23 + b23 + 0x00020002u) >> 2;
                // lanes past the last word repeat the last word's columns (wc is clamped) and store the same value to the
                // same address: no divergent branch around the store, so the unrolled rows are one basic block and the
                // shared-memory loads of the next rows are issued under the arithmetic of the current one
                *reinterpret_cast<uint32_t*>(drow) = __byte_perm(lo, hi, 0x6420);
            }
        }
        __syncthreads();             // Hb / ytile are free for the next block
    }
}

}  // namespace

// resizeMaps.m[l] = level l as a TMA source for the tiled resize of level l + 1 (box of that level).
int orbfe_resize_make_maps(const OrbfeFrameGeom& g, OrbfeChunkBufs& b, int frames) {
    int bw[ORBFE_MAX_LEVELS], bh[ORBFE_MAX_LEVELS];
    size_t smem = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const bool use = l + 1 < g.nlevels && g.lv[l + 1].rzBoxW > 0;
        bw[l] = use ? g.lv[l + 1].rzBoxW : 16;
        bh[l] = use ? g.lv[l + 1].rzBoxH : 1;
        if (use) smem = std::max(smem, 2 * (size_t)((bw[l] * bh[l] + 127) & ~127) + (size_t)bh[l] * ORBFE_RZ_DW * 4);
    }
    if (smem > 200 * 1024) return orbfe_fail(ORBFE_ERR_INVALID, "tiled resize: source box too large", cudaSuccess);
    if (smem > 40 * 1024 &&
        (cudaFuncSetAttribute(k_resize_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
         cudaFuncSetAttribute(k_resize_tile<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess))
        return orbfe_fail(ORBFE_ERR_CUDA, "cudaFuncSetAttribute(k_resize_tile)", cudaGetLastError());
    return orbfe_make_level_maps(g, b.pyr, frames, bw, bh, b.resizeMaps);
}

void orbfe_launch_pyramid(const OrbfeFrameGeom& g, const OrbfeTap* taps, const uint8_t* d_images,
                          size_t step, size_t frameStride, const OrbfeChunkBufs& b, int B,
                          cudaStream_t st, long long* launches, const OrbfeRectify* rect) {
    for (int l = 0; l < g.nlevels; l++) {
        const OrbfeLevelGeom& L = g.lv[l];
        const int total = (L.pitch >> 2) * (L.h + 2 * ORBFE_YOFF);
        dim3 grid((total + 255) / 256, B);
        if (l == 0 && rect && rect->mapx) {
            dim3 g0(((L.pitch >> 2) + 63) / 64, (L.h + 2 * ORBFE_YOFF + 3) / 4, B);
            k_level0_rect<<<g0, 256, 0, st>>>(d_images, step, frameStride, rect->srcRows, rect->srcCols, rect->mapx, rect->mapy,
                                              b.pyr + L.off, g.pyrStride, L.w, L.h, L.pitch);
        } else if (l == 0) {
            const int groups = L.pitch >> 4, H = L.h + 2 * ORBFE_YOFF;
            const int gFirst = (ORBFE_XOFF + 15) / 16;                          // first group with x0 >= 0
            const int gLast = std::min((L.w - 16 + ORBFE_XOFF) / 16, groups - 1); // last group with x0 + 15 < w
            if (L.w >= 16 && gLast >= gFirst) {
                dim3 g0((gLast - gFirst + 1 + 63) / 64, (H + 3) / 4, B);
                k_level0<<<g0, 256, 0, st>>>(d_images, step, frameStride, b.pyr + L.off, g.pyrStride, L.w, L.h, L.pitch, gFirst, gLast);
                const int nb = gFirst + (groups - 1 - gLast);
                if (nb > 0) {
                    k_level0_border<<<dim3((H * nb + 255) / 256, 1, B), 256, 0, st>>>(d_images, step, frameStride, b.pyr + L.off,
                                                                                   g.pyrStride, L.w, L.h, L.pitch, gFirst, gLast);
                    ++*launches;
                }
            } else {   // images narrower than one group: everything is border
                k_level0_border<<<dim3((H * groups + 255) / 256, 1, B), 256, 0, st>>>(d_images, step, frameStride, b.pyr + L.off,
                                                                                   g.pyrStride, L.w, L.h, L.pitch, 0, -1);
            }
        } else {
            const OrbfeLevelGeom& S = g.lv[l - 1];
            const OrbfeTap* xt = taps + L.xtab;
            const OrbfeTap* yt = taps + L.ytab;
            if (L.mode == 0 && L.fastTaps && L.rzBoxW > 0 && !getenv("ORBFE_RESIZE_OLD")) {
                const int nyb = (L.h + 2 * ORBFE_YOFF + ORBFE_RZ_DH - 1) / ORBFE_RZ_DH, nxb = ((L.pitch >> 2) + 31) / 32;
                // several row blocks per CTA (x taps set up once, source box of the next block prefetched) as long as the
                // launch keeps about four waves of CTAs; ORBFE_RZ_PER overrides (tuning)
                static const int perEnv = getenv("ORBFE_RZ_PER") ? atoi(getenv("ORBFE_RZ_PER")) : 0;
                const long long blocks = (long long)nxb * nyb * B;
                const int per = perEnv > 0 ? std::min(perEnv, nyb) : (int)std::max(1LL, std::min((long long)nyb, blocks / (148LL * 24)));
                dim3 gt(nxb, (nyb + per - 1) / per, B);
                const size_t sm = 2 * (size_t)((L.rzBoxW * L.rzBoxH + 127) & ~127) + (size_t)L.rzBoxH * ORBFE_RZ_DW * 4;
                static const int rzWarps = getenv("ORBFE_RZ_WARPS") ? atoi(getenv("ORBFE_RZ_WARPS")) : 4;
                if (rzWarps == 8)
                    k_resize_tile<8><<<gt, 256, sm, st>>>(b.pyr, g.pyrStride, b.resizeMaps.m[l - 1], xt, yt, taps + L.rzXblk,
                                                        taps + L.rzYblk, L.off, L.w, L.h, L.pitch, L.rzBoxW, L.rzBoxH, per);
                else
                    k_resize_tile<4><<<gt, 128, sm, st>>>(b.pyr, g.pyrStride, b.resizeMaps.m[l - 1], xt, yt, taps + L.rzXblk,
                                                        taps + L.rzYblk, L.off, L.w, L.h, L.pitch, L.rzBoxW, L.rzBoxH, per);
            } else if (L.mode == 0 && L.fastTaps) {
                dim3 gf(((L.pitch >> 2) / RS_G + 31) / 32, (L.h + 2 * ORBFE_YOFF + RS_ROWS * RS_WARPS - 1) / (RS_ROWS * RS_WARPS), B);
                k_resize_fast<<<gf, 32 * RS_WARPS, 0, st>>>(b.pyr, g.pyrStride, xt, yt, S.off, S.pitch, L.off, L.w, L.h, L.pitch);
            } else if (L.mode == 0)
                k_resize<0><<<grid, 256, 0, st>>>(b.pyr, g.pyrStride, xt, yt, S.off, S.pitch, L.off, L.w, L.h, L.pitch);
            else if (L.mode == 1)
                k_resize<1><<<grid, 256, 0, st>>>(b.pyr, g.pyrStride, xt, yt, S.off, S.pitch, L.off, L.w, L.h, L.pitch);
            else
                k_resize<2><<<grid, 256, 0, st>>>(b.pyr, g.pyrStride, xt, yt, S.off, S.pitch, L.off, L.w, L.h, L.pitch);
        }
        ++*launches;
    }
}
