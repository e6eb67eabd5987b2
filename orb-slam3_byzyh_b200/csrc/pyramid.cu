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

#include "orbfe_internal.h"
#include "remap_core.h"

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

}  // namespace

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
            if (L.mode == 0 && L.fastTaps) {
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
