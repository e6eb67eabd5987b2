// fast_core.h -- FAST-9/16 arithmetic shared by the FAST kernel and its host unit test.
//
// Restates OpenCV's FAST_t<16> corner test and cornerScore<16> (the primitive the reference
// calls at /root/reference/src/ORBextractor.cc:1135,1144; OpenCV itself is not vendored):
// ring = Bresenham circle r=3, corner at threshold t  <=>  best > t, response = best-1, with
//   best = max over the 16 arcs of 9 contiguous ring pixels of
//          max( min(centre - ring), min(ring - centre) ).
//
// Two pixels are processed per 32-bit register as 16-bit lanes (Blackwell has native 2- and
// 3-input packed min/max: VIMNMX.U16x2 / VIMNMX3.U16x2).  Differences are kept BIASED,
//   e_k = centre + 256 - ring_k  in [1, 511]  per lane,
// so one plain 32-bit subtract serves both lanes (no borrow can cross) and unsigned packed
// min/max order them like the signed differences.  With the 3-wise sliding scheme
//   m3_k = min(e_k, e_k+1, e_k+2),  min9_k = min(m3_k, m3_k+3, m3_k+6)
// an arc minimum costs two 3-input ops.  best = max(0, max_k min9_k - 256, 256 - min_k max9_k).
//
// Four equivalent formulations live here, all checked against the oracle by the host unit test
// (tests/test_fast_core.py) and timed inside k_fast_score on B200 (ms per 1024 frames):
//   fc_margin2          biased differences, 3-wise sliding windows: 80 packed min/max per pixel pair      3.86
//   fc_margin2_raw      the same on the raw ring values (no differences)                                 +2.5 %
//   fc_margin2_pair     arcs taken in pairs (k, k+1 share eight ring pixels): 72 per pixel pair           3.57
//   fc_margin2_pair_raw pairs on the raw ring values (round 1's k_fast_score)                           3.41
//   fc_margin2_pair_raw_biased  the same, no packed subtraction -- THE ONE k_fast_cells USES
//
// NOTE (measured on B200, nvcc 12.9): a formulation that folds `max(best, -mx)` into the running
// maximum is MISCOMPILED for sm_100a (ptxas drops the negation when it fuses into VIMNMX3); this
// formulation negates once, outside the min/max network, and is checked on the device against
// the CPU oracle by tests/test_gpu_extract.py.
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

// ---- packed 16-bit lane helpers (device: DPX/video instructions; host: emulation) ----------
#if defined(__CUDA_ARCH__)
static __device__ __forceinline__ uint32_t fc_min3u(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
static __device__ __forceinline__ uint32_t fc_max3u(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
static __device__ __forceinline__ uint32_t fc_minu(uint32_t a, uint32_t b) { return __vminu2(a, b); }
static __device__ __forceinline__ uint32_t fc_maxu(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
static __device__ __forceinline__ uint32_t fc_add2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
static __device__ __forceinline__ uint32_t fc_sub2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
static __device__ __forceinline__ uint32_t fc_max3s_relu(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2_relu(a, b, c); }
#else
#define FC_LANES(expr_lo, expr_hi) (((uint32_t)(uint16_t)(expr_lo)) | ((uint32_t)(uint16_t)(expr_hi) << 16))
static inline uint32_t fc_lo(uint32_t a) { return a & 0xFFFFu; }
static inline uint32_t fc_hi(uint32_t a) { return a >> 16; }
static inline uint32_t fc_mn(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t fc_mx(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t fc_minu(uint32_t a, uint32_t b) { return FC_LANES(fc_mn(fc_lo(a), fc_lo(b)), fc_mn(fc_hi(a), fc_hi(b))); }
static inline uint32_t fc_maxu(uint32_t a, uint32_t b) { return FC_LANES(fc_mx(fc_lo(a), fc_lo(b)), fc_mx(fc_hi(a), fc_hi(b))); }
static inline uint32_t fc_min3u(uint32_t a, uint32_t b, uint32_t c) { return fc_minu(fc_minu(a, b), c); }
static inline uint32_t fc_max3u(uint32_t a, uint32_t b, uint32_t c) { return fc_maxu(fc_maxu(a, b), c); }
static inline uint32_t fc_add2(uint32_t a, uint32_t b) { return FC_LANES(fc_lo(a) + fc_lo(b), fc_hi(a) + fc_hi(b)); }
static inline uint32_t fc_sub2(uint32_t a, uint32_t b) { return FC_LANES(fc_lo(a) - fc_lo(b), fc_hi(a) - fc_hi(b)); }
static inline int fc_s16(uint32_t v) { return (int)(int16_t)(uint16_t)v; }
static inline int fc_mx3s0(int a, int b, int c) { int m = a > b ? a : b; m = m > c ? m : c; return m > 0 ? m : 0; }
static inline uint32_t fc_max3s_relu(uint32_t a, uint32_t b, uint32_t c) {
    return FC_LANES(fc_mx3s0(fc_s16(fc_lo(a)), fc_s16(fc_lo(b)), fc_s16(fc_lo(c))),
                    fc_mx3s0(fc_s16(fc_hi(a)), fc_s16(fc_hi(b)), fc_s16(fc_hi(c))));
}
#endif

#define FC_BIAS2 0x01000100u  // +256 in both lanes

// e[k] = biased packed differences (centre + 256 - ring_k per lane), k = 0..15 in ring order.
// Returns per lane  max(best - sub, 0)  where `sub2` holds `sub` in both lanes (sub = minThFAST:
// the score map stores the margin over the low threshold, 0 = "not a corner at minThFAST").
// Operand placement: VIMNMX3 reads three registers from a two-bank register file, so an instruction whose operands
// are all fresh needs two fetch cycles; operands that sit in the SAME source slot as in the previous instruction come
// from the operand-reuse cache instead.  Sliding triples (v_k, v_k+1, v_k+2) share two values with their successor:
// keeping value v_j in slot j % 3 makes both reusable (ptxas marks 25 of 32 instead of 13, tools/pipe_bench3.cu).
#define FC_SLOT3(op, v, k, n)                                                                              \
    op(((k) % 3 == 0) ? v[(k) % (n)] : ((k) % 3 == 2) ? v[((k) + 1) % (n)] : v[((k) + 2) % (n)],          \
       ((k) % 3 == 1) ? v[(k) % (n)] : ((k) % 3 == 0) ? v[((k) + 1) % (n)] : v[((k) + 2) % (n)],          \
       ((k) % 3 == 2) ? v[(k) % (n)] : ((k) % 3 == 1) ? v[((k) + 1) % (n)] : v[((k) + 2) % (n)])

static FC_HD uint32_t fc_margin2(const uint32_t* e, uint32_t sub2) {
    uint32_t mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn3[k] = FC_SLOT3(fc_min3u, e, k, 16);
        mx3[k] = FC_SLOT3(fc_max3u, e, k, 16);
    }
    // arcs of 9 = three triples 3 apart: walk k in steps of 3 (3 is coprime with 16), so that consecutive arcs again
    // share two operands; b[i] = m3[3i mod 16] turns the arc (m3_k, m3_k+3, m3_k+6) into the sliding triple (b_i, b_i+1, b_i+2)
    uint32_t bn[16], bx[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { bn[i] = mn3[(3 * i) & 15]; bx[i] = mx3[(3 * i) & 15]; }
    uint32_t A = 0u, B = 0xFFFFFFFFu;  // max_k min9_k, min_k max9_k (biased, unsigned lanes)
#pragma unroll
    for (int i = 0; i < 16; i++) {
        A = fc_maxu(A, FC_SLOT3(fc_min3u, bn, i, 16));
        B = fc_minu(B, FC_SLOT3(fc_max3u, bx, i, 16));
    }
    // best = max(0, A-256, 256-B); margin = max(0, best - sub)
    const uint32_t a = fc_sub2(A, fc_add2(FC_BIAS2, sub2));
    const uint32_t b = fc_sub2(fc_sub2(FC_BIAS2, sub2), B);
    return fc_max3s_relu(a, b, 0u);
}

// The same margin with the arcs taken in PAIRS.  Arcs k and k+1 (k even) share the 8 ring positions k+1..k+8, so
//   max(min9_k, min9_k+1) = min( min(e_k+1..e_k+8), max(e_k, e_k+9) ).
// With p_i = min(e_2i+1, e_2i+2) (8 two-input ops) the shared part is the window p_i..p_i+3 of a circular 8-array:
//   q_i = min3(p_i, p_i+1, p_i+2),  t_i = min3(q_i, p_i+3, max(e_2i, e_2i+9)),  A = max_i t_i
// i.e. 16 two-input + 16 three-input + 4 for the final maximum = 36 packed min/max per polarity instead of 40, and
// 16 of them read two registers instead of three.
static FC_HD void fc_pair_extrema(const uint32_t* e, uint32_t& A, uint32_t& B) {   // A = max_k min9_k, B = min_k max9_k
    uint32_t pn[8], px[8], xn[8], xx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t a = e[2 * i + 1], b = e[(2 * i + 2) & 15];
        pn[i] = fc_minu(a, b);
        px[i] = fc_maxu(a, b);
        const uint32_t c = e[2 * i], d = e[(2 * i + 9) & 15];
        xn[i] = fc_maxu(c, d);   // for the arc minima: the better of the two end pixels
        xx[i] = fc_minu(c, d);   // for the arc maxima
    }
    uint32_t tn[8], tx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t qn = FC_SLOT3(fc_min3u, pn, i, 8);
        const uint32_t qx = FC_SLOT3(fc_max3u, px, i, 8);
        tn[i] = fc_min3u(qn, pn[(i + 3) & 7], xn[i]);
        tx[i] = fc_max3u(qx, px[(i + 3) & 7], xx[i]);
    }
    A = fc_maxu(fc_max3u(fc_max3u(tn[0], tn[1], tn[2]), fc_max3u(tn[3], tn[4], tn[5]), tn[6]), tn[7]);
    B = fc_minu(fc_min3u(fc_min3u(tx[0], tx[1], tx[2]), fc_min3u(tx[3], tx[4], tx[5]), tx[6]), tx[7]);
}
static FC_HD uint32_t fc_margin2_pair(const uint32_t* e, uint32_t sub2) {
    uint32_t A, B;
    fc_pair_extrema(e, A, B);
    const uint32_t a = fc_sub2(A, fc_add2(FC_BIAS2, sub2));
    const uint32_t b = fc_sub2(fc_sub2(FC_BIAS2, sub2), B);
    return fc_max3s_relu(a, b, 0u);
}
// ... and on the raw ring values (see fc_margin2_raw): best = max(0, c - min_k max9_k(ring), max_k min9_k(ring) - c)
static FC_HD uint32_t fc_margin2_pair_raw(const uint32_t* r, uint32_t c2, uint32_t sub2) {
    uint32_t hi, lo;
    fc_pair_extrema(r, hi, lo);
    const uint32_t a = fc_sub2(fc_sub2(c2, sub2), lo);
    const uint32_t b = fc_sub2(hi, fc_add2(c2, sub2));
    return fc_max3s_relu(a, b, 0u);
}

// The same without packed subtractions -- THE ONE k_fast_cells USES.  nvcc 12.9 turns a packed `x - y` into
// x + (~y + 0x10001) and then folds the +1s of several negations together; in one of the three instantiations of
// k_fast_cells that folding came out as `VIADD.16x2 R, ~sub2, 0x0` (the +1 lost: every dark-polarity score one too
// low, found by the parity tests on B200).  Here every lane carries a bias of 1024, so all lanes stay positive and
// plain 32-bit adds / subtracts are exact per lane (no borrow can cross):
//   a = c + 1024 - sub - min_k max9_k  in [514, 1279],   b = max_k min9_k + 1024 - sub - c  in [514, 1279]
//   margin = max(a, b, 1024) - 1024.
static FC_HD uint32_t fc_margin2_pair_raw_biased(const uint32_t* r, uint32_t c2, uint32_t sub2) {
    uint32_t hi, lo;
    fc_pair_extrema(r, hi, lo);
    const uint32_t K = 0x04000400u, ks = K - sub2;   // sub <= 255 per lane: no borrow
    const uint32_t a = c2 + ks - lo;
    const uint32_t b = hi + ks - c2;
    return fc_max3u(a, b, K) - K;
}

// The margin from the RAW ring values (r[k] = two ring pixels as 16-bit lanes, c2 = the two centres), without
// forming the 16 differences:  min over an arc of (c - ring) = c - max over the arc of ring, and
// min over an arc of (ring - c) = (min over the arc of ring) - c, so
//   best = max(0, c - min_k max9_k(ring), max_k min9_k(ring) - c).
// 96 packed min/max per pixel pair plus a 5-instruction tail (the differences cost 16 more).
static FC_HD uint32_t fc_margin2_raw(const uint32_t* r, uint32_t c2, uint32_t sub2) {
    uint32_t mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn3[k] = fc_min3u(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        mx3[k] = fc_max3u(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;  // min_k max9_k, max_k min9_k
#pragma unroll
    for (int k = 0; k < 16; k++) {
        lo = fc_minu(lo, fc_max3u(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]));
        hi = fc_maxu(hi, fc_min3u(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]));
    }
    // lanes are in [0, 255]; the packed subtractions below may go negative per lane (signed 16-bit)
    const uint32_t a = fc_sub2(fc_sub2(c2, sub2), lo);   // (c - sub) - min max9
    const uint32_t b = fc_sub2(hi, fc_add2(c2, sub2));   // max min9 - (c + sub)
    return fc_max3s_relu(a, b, 0u);
}

// ---- dense early reject on packed bytes (four pixels per register) ---------------------------------------------
// Every arc of 9 contiguous ring pixels contains one pixel of each opposite pair {k, k+8} (9 > 8), and all pixels
// of a corner's arc differ from the centre by more than t with one sign.  So a corner at threshold t, of either
// polarity, needs |ring - centre| > t on at least one pixel of EVERY opposite pair; the kernel tests the two compass
// pairs, k = 0/8 (rows y+3 / y-3) and k = 4/12 (columns x+3 / x-3).  Bytes stay packed: the absolute difference is one
// instruction (VABSDIFF4), the comparison with u = t+1 a SWAR byte compare that is exact for every u in [1, 255].
#if defined(__CUDA_ARCH__)
static __device__ __forceinline__ uint32_t fc_absdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
#else
static inline uint32_t fc_absdiff4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
#endif
// Bit 7 of every byte of the result: that byte of a is >= u, with uLow = (u & 0x7F) * 0x01010101 and
// uTop = (u & 0x80) ? ~0 : 0.  (a | 0x80) - low7(u) never borrows across bytes (each byte stays >= 1) and leaves
// "low7(a) >= low7(u)" in bit 7; the top bits decide when they differ.  The other bits are junk.
static FC_HD uint32_t fc_ge4(uint32_t a, uint32_t uLow, uint32_t uTop) {
    const uint32_t z = (a | 0x80808080u) - uLow;
    return (a & ~uTop) | (~(a ^ uTop) & z);
}
// 0x80 in every byte whose pixel passes the compass test (c = four centres, n / s / w / e = the ring pixels
// 3 px above / below / left / right of each of them).  SMALLU: u < 128 is known (thresholds up to 126, i.e. every
// configuration the reference ships): bit 7 of a >= u is then a | z, the OR folds into the pair combination and the
// byte compare costs two instructions instead of three.
template <bool SMALLU>
static FC_HD uint32_t fc_compass4(uint32_t c, uint32_t n, uint32_t s, uint32_t w, uint32_t e, uint32_t uLow, uint32_t uTop) {
    if (SMALLU) {
        const uint32_t an = fc_absdiff4(n, c), as = fc_absdiff4(s, c), aw = fc_absdiff4(w, c), ae = fc_absdiff4(e, c);
        const uint32_t H = 0x80808080u;
        const uint32_t v = an | as | ((an | H) - uLow) | ((as | H) - uLow);
        const uint32_t h = aw | ae | ((aw | H) - uLow) | ((ae | H) - uLow);
        return v & h & H;
    }
    const uint32_t v = fc_ge4(fc_absdiff4(n, c), uLow, uTop) | fc_ge4(fc_absdiff4(s, c), uLow, uTop);
    const uint32_t h = fc_ge4(fc_absdiff4(w, c), uLow, uTop) | fc_ge4(fc_absdiff4(e, c), uLow, uTop);
    return v & h & 0x80808080u;
}

// Scalar convenience (host tests, small kernels): best of one pixel of a byte image.
template <int PITCH>
static FC_HD int fc_best_scalar(const uint8_t* p) {
    uint32_t e[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t r = p[FC_RING_DX(k) + FC_RING_DY(k) * PITCH];
        e[k] = ((uint32_t)p[0] + 256u - r) * 0x00010001u;
    }
    return (int)(fc_margin2(e, 0u) & 0xFFFFu);
}
