// fast_core.h -- FAST-9/16 arithmetic shared by the FAST kernel and its host unit test.
//
// Restates OpenCV's FAST_t<16> corner test and cornerScore<16> (the primitive the reference
// calls at /root/reference/src/ORBextractor.cc:1135,1144; OpenCV itself is not vendored):
// ring = Bresenham circle r=3, corner at threshold t  <=>  best > t, response = best-1, with
//   best = max over the 16 arcs of 9 contiguous ring pixels of
//          max( min(centre - ring), min(ring - centre) ).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FC_HD __host__ __device__ __forceinline__
#else
#define FC_HD inline
#endif

// Ring offsets (dx,dy), OpenCV order.
#define FC_RING_DX(k) ((k) == 0 ? 0 : (k) == 1 ? 1 : (k) == 2 ? 2 : (k) == 3 ? 3 : (k) == 4 ? 3 : (k) == 5 ? 3 : (k) == 6 ? 2 : (k) == 7 ? 1 : (k) == 8 ? 0 : (k) == 9 ? -1 : (k) == 10 ? -2 : (k) == 11 ? -3 : (k) == 12 ? -3 : (k) == 13 ? -3 : (k) == 14 ? -2 : -1)
#define FC_RING_DY(k) ((k) == 0 ? 3 : (k) == 1 ? 3 : (k) == 2 ? 2 : (k) == 3 ? 1 : (k) == 4 ? 0 : (k) == 5 ? -1 : (k) == 6 ? -2 : (k) == 7 ? -3 : (k) == 8 ? -3 : (k) == 9 ? -3 : (k) == 10 ? -2 : (k) == 11 ? -1 : (k) == 12 ? 0 : (k) == 13 ? 1 : (k) == 14 ? 2 : 3)

static FC_HD int fc_min(int a, int b) { return a < b ? a : b; }
static FC_HD int fc_max(int a, int b) { return a > b ? a : b; }

// Quick reject (OpenCV's pair test): a 9-arc contains one pixel of every opposite ring pair,
// so each pair needs a member darker than v-t (bit 0) or brighter than v+t (bit 1).
template <int PITCH>
static FC_HD bool fc_may_be_corner(const uint8_t* p, int th) {
    const int v = p[0], lo = v - th, hi = v + th;
    int d = 3;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // visit the pairs in OpenCV's order 0,4,2,6,1,3,5,7 so flat regions exit early
        const int kk = (k == 0) ? 0 : (k == 1) ? 4 : (k == 2) ? 2 : (k == 3) ? 6 : (k == 4) ? 1 : (k == 5) ? 3 : (k == 6) ? 5 : 7;
        const int a = p[FC_RING_DX(kk) + FC_RING_DY(kk) * PITCH];
        const int b = p[FC_RING_DX(kk + 8) + FC_RING_DY(kk + 8) * PITCH];
        d &= ((a < lo ? 1 : 0) | (a > hi ? 2 : 0)) | ((b < lo ? 1 : 0) | (b > hi ? 2 : 0));
        if (d == 0) return false;
    }
    return true;
}

// best (see header comment); sliding 9-window min / max over the circular 16-ring by doubling.
template <int PITCH>
static FC_HD int fc_arc_best(const uint8_t* p) {
    int d[16];
    const int v = p[0];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - (int)p[FC_RING_DX(k) + FC_RING_DY(k) * PITCH];
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn2[k] = fc_min(d[k], d[(k + 1) & 15]);
        mx2[k] = fc_max(d[k], d[(k + 1) & 15]);
    }
    int mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn4[k] = fc_min(mn2[k], mn2[(k + 2) & 15]);
        mx4[k] = fc_max(mx2[k], mx2[(k + 2) & 15]);
    }
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int mn9 = fc_min(fc_min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx9 = fc_max(fc_max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = fc_max(best, fc_max(mn9, -mx9));
    }
    return best;
}
