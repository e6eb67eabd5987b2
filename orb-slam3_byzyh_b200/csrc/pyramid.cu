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
#include "orbfe_internal.h"

namespace {

__device__ __forceinline__ int reflect101_clamped(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * (len - 1) - p;
    return max(0, min(p, len - 1));
}

__global__ void __launch_bounds__(256)
k_level0(const uint8_t* __restrict__ img, size_t step, size_t frameStride, uint8_t* __restrict__ pyr,
         unsigned long long pyrStride, int w, int h, int pitch) {
    const int words = pitch >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int H = h + 2 * ORBFE_YOFF;
    if (idx >= words * H) return;
    const int py = idx / words, wc = idx - py * words;
    const uint8_t* src = img + (size_t)blockIdx.y * frameStride +
                         (size_t)reflect101_clamped(py - ORBFE_YOFF, h) * step;
    const int x0 = 4 * wc - ORBFE_XOFF;
    uint32_t out;
    if (x0 >= 0 && x0 + 3 < w && ((reinterpret_cast<size_t>(src + x0) & 3) == 0)) {
        out = __ldg(reinterpret_cast<const uint32_t*>(src + x0));
    } else {
        out = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            out |= (uint32_t)__ldg(src + reflect101_clamped(x0 + i, w)) << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(pyr + (size_t)blockIdx.y * pyrStride + (size_t)py * pitch + 4 * wc) = out;
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
// owns four adjacent PADDED destination columns and walks down RS_ROWS destination rows: the x-taps
// live in registers, every source row segment is fetched as three aligned 32-bit words, one PRMT
// per pixel pulls its two tap bytes out of that 8-byte window and one IDP2A applies the two 11-bit
// weights.  The horizontal pass of a source row is kept (already >> 4) and reused when the next
// destination row needs the same source row (at scale 1.2 that is 5 rows out of 6).  Border
// columns/rows are the same computation on the reflected destination coordinate, so resize +
// copyMakeBorder stay one pass with coalesced 32-bit stores.
constexpr int RS_ROWS = 16, RS_WARPS = 4;

__global__ void __launch_bounds__(32 * RS_WARPS)
k_resize_fast(uint8_t* pyr, unsigned long long pyrStride, const OrbfeTap* __restrict__ xtab,
              const OrbfeTap* __restrict__ ytab, unsigned srcOff, int srcPitch, unsigned dstOff, int w, int h,
              int pitch) {
    const int words = pitch >> 2;
    const bool active = blockIdx.x * 32 + (int)(threadIdx.x & 31) < words;   // no early exit: the warp shuffles below
    const int wc = min(blockIdx.x * 32 + (int)(threadIdx.x & 31), words - 1);
    const int H = h + 2 * ORBFE_YOFF;
    const int py0 = (blockIdx.y * RS_WARPS + (threadIdx.x >> 5)) * RS_ROWS;
    if (py0 >= H) return;
    uint8_t* base = pyr + (size_t)blockIdx.z * pyrStride;
    int lo = 1 << 30;
    OrbfeTap tp[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        tp[i] = xtab[reflect101_clamped(4 * wc - ORBFE_XOFF + i, w)];
        lo = min(lo, (int)tp[i].s);
    }
    uint32_t wgt[4], sel[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        wgt[i] = (uint32_t)(uint16_t)tp[i].a0 | ((uint32_t)(uint16_t)tp[i].a1 << 16);
        sel[i] = (uint32_t)(tp[i].s - lo) | ((uint32_t)(tp[i].s1 - lo) << 4);   // byte0 = src[s], byte1 = src[s1]
    }
    const int col = ORBFE_XOFF + lo, sh = 8 * (col & 3);
    const int srcWords = srcPitch >> 2;
    const int wi0 = col >> 2, wi1 = min(wi0 + 1, srcWords - 1), wi2 = min(wi0 + 2, srcWords - 1);
    const uint32_t* __restrict__ sroi = reinterpret_cast<const uint32_t*>(base + srcOff + (size_t)ORBFE_YOFF * srcPitch);

    // horizontal pass of one source row segment (three words), already >> 4
    auto hpass = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t* Hs) {
        const uint32_t lo8 = __funnelshift_r(w0, w1, sh), hi8 = __funnelshift_r(w1, w2, sh);
#pragma unroll
        for (int i = 0; i < 4; i++) Hs[i] = __dp2a_lo(wgt[i], __byte_perm(lo8, hi8, sel[i]), 0u) >> 4;
    };

    // The y-taps of the strip are fetched once (lane i holds the tap of row py0+i) and handed out
    // by shuffle; the source words of the NEXT destination row are requested before the current row
    // is computed, so no global-load latency sits on the row-to-row dependency chain.
    const int lane = threadIdx.x & 31;
    OrbfeTap myTap = {0, 0, 0, 0};
    if (lane < RS_ROWS) myTap = ytab[reflect101_clamped(min(py0 + lane, H - 1) - ORBFE_YOFF, h)];
    const unsigned tapS = (uint32_t)(uint16_t)myTap.s | ((uint32_t)(uint16_t)myTap.s1 << 16);
    const unsigned tapA = (uint32_t)(uint16_t)myTap.a0 | ((uint32_t)(uint16_t)myTap.a1 << 16);

    int c0 = -1, c1 = -1;
    uint32_t C0[4] = {0, 0, 0, 0}, C1[4] = {0, 0, 0, 0};   // cached horizontal passes of source rows c0, c1
    const int n = min(RS_ROWS, H - py0);
    uint32_t* dst = reinterpret_cast<uint32_t*>(base + dstOff) + wc + (size_t)py0 * words;
    unsigned rs = __shfl_sync(0xffffffffu, tapS, 0);
    uint32_t p0[3], p1[3];   // prefetched words of source rows r0 and r1 of the current row
    {
        const uint32_t* ra = sroi + (size_t)(rs & 0xFFFFu) * srcWords;
        const uint32_t* rb = sroi + (size_t)(rs >> 16) * srcWords;
        p0[0] = ra[wi0]; p0[1] = ra[wi1]; p0[2] = ra[wi2];
        p1[0] = rb[wi0]; p1[1] = rb[wi1]; p1[2] = rb[wi2];
    }
    for (int i = 0; i < n; i++, dst += words) {
        const int r0 = (int)(rs & 0xFFFFu), r1 = (int)(rs >> 16);
        const unsigned wa = __shfl_sync(0xffffffffu, tapA, i);
        uint32_t q0[3] = {p0[0], p0[1], p0[2]}, q1[3] = {p1[0], p1[1], p1[2]};
        if (i + 1 < n) {   // request the next row's source words now
            rs = __shfl_sync(0xffffffffu, tapS, i + 1);
            const uint32_t* ra = sroi + (size_t)(rs & 0xFFFFu) * srcWords;
            const uint32_t* rb = sroi + (size_t)(rs >> 16) * srcWords;
            p0[0] = ra[wi0]; p0[1] = ra[wi1]; p0[2] = ra[wi2];
            p1[0] = rb[wi0]; p1[1] = rb[wi1]; p1[2] = rb[wi2];
        }
        // all threads of a warp share the row, hence r0/r1/c0/c1: the branches below are warp-uniform
        if (r0 != c0) {
            if (r0 == c1) {
#pragma unroll
                for (int k = 0; k < 4; k++) C0[k] = C1[k];
            } else {
                hpass(q0[0], q0[1], q0[2], C0);
            }
            c0 = r0;
        }
        if (r1 != c1) {
            if (r1 == c0) {
#pragma unroll
                for (int k = 0; k < 4; k++) C1[k] = C0[k];
            } else {
                hpass(q1[0], q1[1], q1[2], C1);
            }
            c1 = r1;
        }
        // ((b*(H>>4))>>16) == umulhi(b << 16, H >> 4) for the non-negative 11-bit weights
        const uint32_t b0 = wa << 16, b1 = wa & 0xFFFF0000u;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (__umulhi(b0, C0[k]) + __umulhi(b1, C1[k]) + 2u) >> 2;
        if (active) *dst = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
    }
}

}  // namespace

void orbfe_launch_pyramid(const OrbfeFrameGeom& g, const OrbfeTap* taps, const uint8_t* d_images,
                          size_t step, size_t frameStride, const OrbfeChunkBufs& b, int B,
                          cudaStream_t st, long long* launches) {
    for (int l = 0; l < g.nlevels; l++) {
        const OrbfeLevelGeom& L = g.lv[l];
        const int total = (L.pitch >> 2) * (L.h + 2 * ORBFE_YOFF);
        dim3 grid((total + 255) / 256, B);
        if (l == 0) {
            k_level0<<<grid, 256, 0, st>>>(d_images, step, frameStride, b.pyr + L.off, g.pyrStride,
                                           L.w, L.h, L.pitch);
        } else {
            const OrbfeLevelGeom& S = g.lv[l - 1];
            const OrbfeTap* xt = taps + L.xtab;
            const OrbfeTap* yt = taps + L.ytab;
            if (L.mode == 0 && L.fastTaps) {
                dim3 gf(((L.pitch >> 2) + 31) / 32, (L.h + 2 * ORBFE_YOFF + RS_ROWS * RS_WARPS - 1) / (RS_ROWS * RS_WARPS), B);
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
