// fast.cu -- the FAST half of ORBextractor::ComputeKeyPointsOctTree
// (/root/reference/src/ORBextractor.cc:1061-1165): per 35-px cell cv::FAST(iniThFAST, nms) with
// the minThFAST retry when the cell came back empty, candidates emitted cell row-major and
// FAST row-major inside a cell.
//
// B200 design.  OpenCV's FAST response is threshold independent: with
//   best(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min(c-ring), min(ring-c))
// a pixel is a corner at threshold t iff best > t and its response is best-1.  The 700 tiny
// per-cell cv::FAST calls of the reference therefore collapse into three streaming kernels:
//   k_fast_score : one pass per pyramid level writing the margin max(best - subTh, 0) of every
//                  pixel of the level's FAST domain as u16 (subTh = max(min(iniThFAST, minThFAST), 1));
//   k_fast_nms   : 3x3 non-max suppression with the reference's per-cell semantics (a neighbour
//                  outside the pixel's own cell interior counts as 0) on 16-bit lanes, two pixels
//                  per instruction; writes two 1-bit-per-pixel maps: survivors at minThFAST and
//                  survivors at iniThFAST;
//   k_fast_cells : one warp per cell (lane = row): popcounts decide "empty at iniThFAST -> use
//                  minThFAST" (:1141-1148), a warp scan gives every row its output rank, the set
//                  bits are emitted in row-major order (= cv::FAST's order).
// NMS is threshold independent: a non-corner neighbour has best_n <= t < best, so comparing
// against its true margin instead of 0 never changes the outcome; cell interiors tile the FAST
// domain exactly (cell j owns columns [19 + j*wCell, 19 + (j+1)*wCell)), so "outside the cell"
// is a per-column / per-row mask.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fast_core.h"
#include "tma.h"
#include "octree_core.h"  // OC_PACK
#include "orbfe_internal.h"

namespace {

// ---- k_fast_score -------------------------------------------------------------------------------
// Tile = 128 x 16 output pixels (+3 halo).  The tile is staged in shared memory already widened to
// 16-bit lanes, in four copies shifted by 0..3 pixels, so that ANY run of four horizontally
// adjacent pixels is one aligned LDS.64 (two u16x2 pairs) -- the 16 ring operands of four pixels
// cost 16 LDS.64 and no byte-extraction ALU work, leaving the ALU pipe to the packed min/max
// network of fast_core.h (2 pixels per instruction).  Each thread scores 4 pixels on 2 rows.
constexpr int TW = ORBFE_FAST_TW, TH = ORBFE_FAST_TH;
constexpr int TROWS = TH + 6;         // 22 staged rows
constexpr int TG = TW / 4 + 1;        // 33 four-pixel groups per staged row and copy
static_assert(TW == 128 && TH == 16, "thread mapping below assumes 128x16 tiles");
static_assert(3 * 8 >= TH + 6, "three row copies per warp cover the staged rows");

// A CTA walks `tilesPerCta` consecutive tiles of one frame and keeps the raw bytes of the NEXT TWO tiles in flight:
// a staged tile (22 rows x 144 bytes of the padded level) is one TMA tensor copy (cp.async.bulk.tensor.3d global ->
// shared through the level's CUtensorMap, issued by a single thread, completion counted in bytes on an mbarrier) into
// one of two raw staging buffers.  No thread holds a register or issues a load for the tile data, and the global-load
// latency hides behind the min/max network instead of stalling every warp of the CTA at its start (measured, 1024
// frames: 4.49 ms with one tile per CTA and plain loads; tools/fast_per_sweep.sh sweeps the tiles per CTA).
// Rows / columns of the box that fall outside the padded level are zero-filled by the TMA unit; they never reach the
// ring of a scored pixel (the FAST domain ends 19 px inside the ROI, the ring reaches 3 px, the border is 19 px).
constexpr int RAWW = 36;                          // words per staged row: 144 bytes = 33 groups + the shifted-in word, 16-byte multiple
constexpr uint32_t RAW_ROW_BYTES = 4 * RAWW;
constexpr uint32_t RAW_TILE_BYTES = RAW_ROW_BYTES * TROWS;

// Thread 0 requests tile t of frame `frame` into `raw`.
__device__ __forceinline__ void fast_request(const OrbfeFrameGeom& g, const OrbfeFastMaps& maps, int frame, int t, void* raw,
                                             uint64_t* bar) {
    if (threadIdx.x != 0) return;
    int l = 0;
    while (l + 1 < g.nlevels && t >= g.lv[l + 1].fastTileBase) l++;
    const OrbfeLevelGeom& L = g.lv[l];
    const int tl = t - L.fastTileBase;
    const int ty = tl / L.fastTilesX, tx = tl - ty * L.fastTilesX;
    mbar_expect_tx(bar, RAW_TILE_BYTES);
    // FAST domain origin = ROI (19,19).  Staged column 0 = ROI x 16 + 128*tx = padded byte column
    // 48 + 128*tx (16-byte aligned); staged row 0 = ROI y 16 + 16*ty = padded row 35 + 16*ty.
    tma_tile_g2s(raw, &maps.m[l], ORBFE_XOFF + 16 + TW * tx, ORBFE_YOFF + 16 + TH * ty, frame, bar);
}

__global__ void __launch_bounds__(256)
k_fast_score(const __grid_constant__ OrbfeFrameGeom g, const __grid_constant__ OrbfeFastMaps maps,
             uint16_t* __restrict__ score, int tilesPerCta) {
    __shared__ __align__(16) uint2 cp[4][TROWS][TG];
    struct __align__(128) RawTile { uint32_t w[TROWS][RAWW]; };   // TMA destinations are 128-byte aligned: sizeof = 3200
    __shared__ RawTile raw[2];
    __shared__ __align__(8) uint64_t bar[2];
    const size_t frameOff = (size_t)blockIdx.y * g.pyrStride;
    const int t0 = blockIdx.x * tilesPerCta, t1 = min(t0 + tilesPerCta, g.fastTiles);
    const uint32_t sub2 = (uint32_t)g.subTh * 0x00010001u;
    const int gq = threadIdx.x & 31, rp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init_fence();
    }
    __syncthreads();
    fast_request(g, maps, blockIdx.y, t0, &raw[0], &bar[0]);      // two tiles in flight
    if (t0 + 1 < t1) fast_request(g, maps, blockIdx.y, t0 + 1, &raw[1], &bar[1]);
    for (int t = t0; t < t1; t++) {
        const int b = (t - t0) & 1;
        const uint32_t parity = ((t - t0) >> 1) & 1;
        mbar_wait(&bar[b], parity);                      // tile t has landed in raw[b]
        if (t > t0) __syncthreads();                     // every warp is done reading cp of the previous tile
        for (int i = threadIdx.x; i < TROWS * TG; i += 256) {
            const int r = i / TG, q = i - r * TG;
            const uint32_t a = raw[b].w[r][q], c = raw[b].w[r][q + 1];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint32_t v = s ? __funnelshift_r(a, c, 8 * s) : a;   // pixels 4q+s .. 4q+s+3
                cp[s][r][q] = make_uint2(__byte_perm(v, 0u, 0x4140), __byte_perm(v, 0u, 0x4342));
            }
        }
        __syncthreads();                                 // cp is complete; raw[b] is free again
        if (t + 2 < t1) {
            // the generic-proxy reads of raw[b] above are ordered before the async-proxy writes of the next request
            fence_proxy_async();
            fast_request(g, maps, blockIdx.y, t + 2, &raw[b], &bar[b]);
        }
        int l = 0;
        while (l + 1 < g.nlevels && t >= g.lv[l + 1].fastTileBase) l++;
        const OrbfeLevelGeom& L = g.lv[l];
        const int tl = t - L.fastTileBase, ty = tl / L.fastTilesX, tx = tl - ty * L.fastTilesX;
        const int x = 19 + TW * tx + 4 * gq;       // ROI x of the first of this thread's 4 pixels
        if (x >= L.w - 19) continue;
        // score map column = ROI x + 13, so that a 4-pixel group is one aligned 64-bit store
        uint16_t* dst = score + frameOff + L.off + ORBFE_SXOFF + x;
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const int orow = 2 * rp + rr, y = 19 + TH * ty + orow;
            if (y >= L.h - 19) break;
            const uint2 c = cp[3][orow + 3][gq];
            // raw ring values: the network works on them directly, two pixels per register (fast_core.h)
            uint32_t r0[16], r1[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int o = 3 + FC_RING_DX(k);
                const uint2 v = cp[o & 3][orow + 3 + FC_RING_DY(k)][gq + (o >> 2)];
                r0[k] = v.x;
                r1[k] = v.y;
            }
            const uint32_t m0 = fc_margin2_pair_raw(r0, c.x, sub2), m1 = fc_margin2_pair_raw(r1, c.y, sub2);
            *reinterpret_cast<uint2*>(dst + (size_t)(ORBFE_YOFF + y) * L.pitch) = make_uint2(m0, m1);
        }
    }
}

// ---- k_fast_nms ---------------------------------------------------------------------------------
// A warp owns a strip of 128 columns x NMS_ROWS rows of the FAST domain; lane = 4 adjacent pixels
// held as two u16x2 registers; the strip is walked top to bottom with a 3-row window in registers.
constexpr int NMS_ROWS = 16, NMS_WARPS = 8;   // CTA = 128 x 128 pixels

struct NmsRow {   // per row, per thread: the horizontal 3-max with and without the centre
    uint32_t A, B;          // centre pairs (p0,p1), (p2,p3)
    uint32_t fullA, fullB;  // max(left, centre, right)
    uint32_t lrA, lrB;      // max(left, right)
};

struct NmsMasks { uint32_t lA, rA, lB, rB; };

struct NmsRaw { uint2 c; uint32_t el, er; };   // a lane's 4 pixels + the words beyond the warp's edges

// Issues the loads of one score-map row segment (4 pixels per lane); nothing is consumed here, so
// the caller can request row y+2 before it works on row y+1.
__device__ __forceinline__ NmsRaw nms_fetch_row(const uint16_t* __restrict__ row, bool colIn, bool needL, bool needR) {
    NmsRaw r;
    r.c = make_uint2(0u, 0u);
    r.el = 0u; r.er = 0u;
    if (colIn) r.c = *reinterpret_cast<const uint2*>(row);
    if (needL) r.el = *reinterpret_cast<const uint32_t*>(row - 2);     // lane 0: pixels left of the warp
    if (needR) r.er = *reinterpret_cast<const uint32_t*>(row + 4);     // lane 31: pixels right of the warp
    return r;
}

// Forms the horizontal maxima of a fetched row.
__device__ __forceinline__ void nms_finish_row(const NmsRaw& raw, int lane, const NmsMasks& m, NmsRow& R) {
    const uint2 c = raw.c;
    uint32_t lw = __shfl_up_sync(0xffffffffu, c.y, 1), rw = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lane == 0) lw = raw.el;
    if (lane == 31) rw = raw.er;
    const uint32_t s1 = __funnelshift_r(c.x, c.y, 16);               // (p1, p2)
    const uint32_t s0 = __funnelshift_r(lw, c.x, 16) & m.lA;          // (p-1, p0): left neighbours of A
    const uint32_t s2 = __funnelshift_r(c.y, rw, 16) & m.rB;          // (p3, p4): right neighbours of B
    const uint32_t s1a = s1 & m.rA, s1b = s1 & m.lB;
    R.A = c.x; R.B = c.y;
    R.lrA = fc_maxu(s0, s1a); R.lrB = fc_maxu(s1b, s2);
    R.fullA = fc_max3u(s0, s1a, c.x); R.fullB = fc_max3u(s1b, s2, c.y);
}

// One output row: 3x3 strict maximum test of `cur` against its eight neighbours.
__device__ __forceinline__ void nms_emit_row(const NmsRow& up, const NmsRow& cur, const NmsRow& dn, bool rowFirst,
                                             bool rowLast, uint32_t ini2, uint32_t min2, unsigned grp, int nibShift, uint32_t nibMask, bool writer,
                                             uint32_t* __restrict__ oMin, uint32_t* __restrict__ oIni) {
    const uint32_t uA = rowFirst ? 0u : up.fullA, uB = rowFirst ? 0u : up.fullB;
    const uint32_t dA = rowLast ? 0u : dn.fullA, dB = rowLast ? 0u : dn.fullB;
    const uint32_t nbA = fc_max3u(uA, dA, cur.lrA), nbB = fc_max3u(uB, dB, cur.lrB);
    // bit 15 of a lane of (x | 0x8000) - a is 0  <=>  a > x   (values < 2^15: no cross-lane borrow)
    uint32_t kA = (nbA | 0x80008000u) - cur.A, kB = (nbB | 0x80008000u) - cur.B;
    const uint32_t iA = (ini2 - cur.A) | kA, iB = (ini2 - cur.B) | kB;    // bit 15 set = NOT kept
    if (min2 != 0x80008000u) { kA |= min2 - cur.A; kB |= min2 - cur.B; }  // warp-uniform, normally skipped
    // gather the four "not kept" flags (bits 15/31 of A, 15/31 of B) into a nibble, then invert
    const uint32_t fMin = __byte_perm(kA, kB, 0x7531) & 0x80808080u;   // bytes: A.b1, A.b3, B.b1, B.b3
    const uint32_t fIni = __byte_perm(iA, iB, 0x7531) & 0x80808080u;
    // multiply packs the four flag bits (7,15,23,31) into bits 28..31: (f >> 7) * 0x10204080 >> 28
    const uint32_t nMin = (~(((fMin >> 7) * 0x10204080u) >> 28)) & nibMask;
    const uint32_t nIni = (~(((fIni >> 7) * 0x10204080u) >> 28)) & nibMask;
    // OR the nibbles of 8 adjacent lanes into one 32-bit word per bitmap with three shuffles: both
    // bitmaps ride in one register (min nibbles in the low half, ini nibbles in the high half) for
    // the two steps inside a 4-lane group, the third step joins the two 16-bit halves.
    uint32_t v = (nMin << nibShift) | (nIni << (nibShift + 16));      // nibShift = 4 * (lane & 3)
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    const uint32_t o = __shfl_xor_sync(0xffffffffu, v, 4);            // the other 4-lane group of the octet
    const uint32_t wMin = __byte_perm(v, o, 0x5410), wIni = __byte_perm(v, o, 0x7632);   // valid on lanes 8k..8k+3
    if (writer) { *oMin = wMin; *oIni = wIni; }
}

__global__ void __launch_bounds__(32 * NMS_WARPS)
k_fast_nms(const __grid_constant__ OrbfeFrameGeom g, const uint16_t* __restrict__ score,
           uint32_t* __restrict__ bits) {
    int l = 0;
    const int t = blockIdx.x;
    while (l + 1 < g.nlevels && t >= g.lv[l + 1].nmsTileBase) l++;
    const OrbfeLevelGeom& L = g.lv[l];
    const int tl = t - L.nmsTileBase;
    const int ty = tl / L.nmsTilesX, tx = tl - ty * L.nmsTilesX;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int dw = L.w - 38, dh = L.h - 38;                 // FAST domain size
    const int dx0 = 128 * tx + 4 * lane;                    // domain x of this thread's first pixel
    const int dy0 = (NMS_WARPS * ty + wid) * NMS_ROWS;      // first domain row of this warp's strip
    if (dy0 >= dh) return;
    const size_t fo = (size_t)blockIdx.y * g.pyrStride + L.off;
    const int pitch = L.pitch;
    // element (domain x, domain y) lives at S[y*pitch + x]
    const uint16_t* row = score + fo + (size_t)(ORBFE_YOFF + 19 + dy0 - 1) * pitch + ORBFE_SXOFF + 19 + dx0;
    const size_t bo = (size_t)blockIdx.y * g.bmWordsPerFrame + (size_t)dy0 * L.bmPitch + 4 * tx + (lane >> 3);
    uint32_t* oMin = bits + bo + L.bmMin;
    uint32_t* oIni = bits + bo + L.bmIni;

    // column masks: a neighbour on the other side of a cell-interior boundary (or outside the
    // domain) counts as 0.  first(x) <=> x % wCell == 0, last(x) <=> x % wCell == wCell-1 || x >= dw-1
    NmsMasks m;
    {
        const int r0 = dx0 % L.wCell;
        uint32_t f = 0, la = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int r = r0 + i;
            if (r >= L.wCell) r -= L.wCell;
            if (r == 0) f |= 1u << i;
            if (r == L.wCell - 1 || dx0 + i >= dw - 1) la |= 1u << i;
        }
        m.lA = ((f & 1) ? 0u : 0xFFFFu) | ((f & 2) ? 0u : 0xFFFF0000u);
        m.lB = ((f & 4) ? 0u : 0xFFFFu) | ((f & 8) ? 0u : 0xFFFF0000u);
        m.rA = ((la & 1) ? 0u : 0xFFFFu) | ((la & 2) ? 0u : 0xFFFF0000u);
        m.rB = ((la & 4) ? 0u : 0xFFFFu) | ((la & 8) ? 0u : 0xFFFF0000u);
    }
    const bool colIn = dx0 < dw;           // this thread's group starts inside the domain
    const bool needL = lane == 0 && colIn && dx0 > 0, needR = lane == 31 && dx0 + 4 < dw;
    const bool writer = (lane & 7) == 0;
    const uint32_t ini2 = ((uint32_t)max(g.iniTh - g.subTh, 0) * 0x00010001u) | 0x80008000u;
    // minThFAST above iniThFAST (unusual, but the reference just calls cv::FAST with it): the retry map
    // then needs its own margin test as well; 0 in the usual case iniThFAST >= minThFAST
    const uint32_t min2 = ((uint32_t)max(g.minTh - g.subTh, 0) * 0x00010001u) | 0x80008000u;
    const unsigned grp = 0xFFu << (lane & 24);
    const int nibShift = 4 * (lane & 3);
    const uint32_t nibMask = colIn ? 0xFu : 0u;            // lanes outside the domain add nothing

    // Rows dy0-1 and dy0+n are only read as neighbours and masked by rowFirst/rowLast at the domain
    // edge; they exist in memory because the level keeps its 19-px border rows.
    NmsRow r0, r1, r2;
    {
        const NmsRaw a0 = nms_fetch_row(row, colIn, needL, needR);
        const NmsRaw a1 = nms_fetch_row(row + pitch, colIn, needL, needR);
        nms_finish_row(a0, lane, m, r0);
        nms_finish_row(a1, lane, m, r1);
    }
    row += 2 * pitch;                       // next row to fetch: dy0 + 1
    const int n = min(NMS_ROWS, dh - dy0);
    int ry = dy0 % L.hCell;                 // row index inside the cell interior
    const int hLast = L.hCell - 1, bmPitch = L.bmPitch;
    NmsRaw nxt = nms_fetch_row(row, colIn, needL, needR);
    // three rows per iteration so that the 3-row window rotates without register moves; the loads
    // of the row after next are always in flight while a row is being processed
    for (int i = 0; i < n; i += 3) {
#define NMS_STEP(UP, CUR, DN, K)                                                                                  \
        if (i + K < n) {                                                                                          \
            const NmsRaw got = nxt;                                                                               \
            row += pitch;                                                                                         \
            if (i + K + 1 < n) nxt = nms_fetch_row(row, colIn, needL, needR);                                     \
            nms_finish_row(got, lane, m, DN);                                                                     \
            nms_emit_row(UP, CUR, DN, ry == 0, ry == hLast || dy0 + i + K == dh - 1, ini2, min2, grp, nibShift,     \
                         nibMask,                                                                                 \
                         writer, oMin, oIni);                                                                     \
            oMin += bmPitch; oIni += bmPitch;                                                                     \
            ry = ry == hLast ? 0 : ry + 1;                                                                        \
        }
        NMS_STEP(r0, r1, r2, 0)
        NMS_STEP(r1, r2, r0, 1)
        NMS_STEP(r2, r0, r1, 2)
#undef NMS_STEP
    }
}

// ---- k_fast_cells -------------------------------------------------------------------------------
__device__ __forceinline__ void row_bits(const uint32_t* __restrict__ row, int pitchWords, int b0, int nb,
                                         uint32_t& r0, uint32_t& r1, uint32_t& r2) {
    const int k = b0 >> 5, sh = b0 & 31;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = k + i < pitchWords ? row[k + i] : 0u;
    r0 = __funnelshift_r(w[0], w[1], sh);
    r1 = __funnelshift_r(w[1], w[2], sh);
    r2 = __funnelshift_r(w[2], w[3], sh);
    if (nb < 32) { r0 &= (1u << nb) - 1u; r1 = 0u; r2 = 0u; }
    else if (nb < 64) { r1 &= (1u << (nb - 32)) - 1u; r2 = 0u; }
    else if (nb < 96) { r2 &= (1u << (nb - 64)) - 1u; }
}

__global__ void __launch_bounds__(256)
k_fast_cells(const __grid_constant__ OrbfeFrameGeom g, const uint16_t* __restrict__ score,
             const uint32_t* __restrict__ bits, uint32_t* __restrict__ slots, int* __restrict__ cellCount) {
    const int lane = threadIdx.x & 31;
    const int cell = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (cell >= g.cellsPerFrame) return;
    int l = 0;
    while (l + 1 < g.nlevels && cell >= g.lv[l + 1].cellBase) l++;
    const OrbfeLevelGeom& L = g.lv[l];
    const int ci = cell - L.cellBase;
    const int i = ci / L.nCols, j = ci - i * L.nCols;
    int* cnt = cellCount + (size_t)blockIdx.y * g.cellsPerFrame + cell;
    // cell geometry, ORBextractor.cc:1098-1132
    const int iniY = ORBFE_FAST_BORDER + i * L.hCell, iniX = ORBFE_FAST_BORDER + j * L.wCell;
    const int maxY = min(iniY + L.hCell + 6, L.maxBY), maxX = min(iniX + L.wCell + 6, L.maxBX);
    if (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6 || maxX - iniX < 7 || maxY - iniY < 7) {
        if (lane == 0) *cnt = 0;
        return;
    }
    const int x0 = iniX + 3, x1 = maxX - 3, y0 = iniY + 3, y1 = maxY - 3;  // FAST interior (ROI coords)
    const int nbits = x1 - x0, b0 = x0 - 19;
    const uint32_t* bmMin = bits + (size_t)blockIdx.y * g.bmWordsPerFrame + L.bmMin;
    const uint32_t* bmIni = bits + (size_t)blockIdx.y * g.bmWordsPerFrame + L.bmIni;
    // phase 1: is the cell empty at iniThFAST?
    int nIni = 0;
    for (int yb = y0; yb < y1; yb += 32) {
        const int y = yb + lane;
        uint32_t r0 = 0, r1 = 0, r2 = 0;
        if (y < y1) row_bits(bmIni + (size_t)(y - 19) * L.bmPitch, L.bmPitch, b0, nbits, r0, r1, r2);
        nIni += __popc(r0) + __popc(r1) + __popc(r2);
    }
    nIni = __reduce_add_sync(0xffffffffu, nIni);
    const uint32_t* bm = nIni > 0 ? bmIni : bmMin;   // :1141-1148 retry at minThFAST only when empty
    // phase 2: ordered emission
    const uint16_t* S = score + (size_t)blockIdx.y * g.pyrStride + L.off + (size_t)ORBFE_YOFF * L.pitch + ORBFE_SXOFF;
    uint32_t* out = slots + (size_t)blockIdx.y * g.slotsPerFrame + L.slotBase + (size_t)ci * L.cellCap;
    int base = 0;
    for (int yb = y0; yb < y1; yb += 32) {
        const int y = yb + lane;
        uint32_t r[3] = {0u, 0u, 0u};
        if (y < y1) row_bits(bm + (size_t)(y - 19) * L.bmPitch, L.bmPitch, b0, nbits, r[0], r[1], r[2]);
        const int c = __popc(r[0]) + __popc(r[1]) + __popc(r[2]);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int pos = base + incl - c;
        base += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t m = r[k];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int x = x0 + 32 * k + b;
                const int margin = S[(size_t)y * L.pitch + x];
                if (pos < L.cellCap)
                    out[pos] = OC_PACK(x - ORBFE_FAST_BORDER, y - ORBFE_FAST_BORDER, margin + g.subTh - 1);
                pos++;
            }
        }
    }
    if (lane == 0) *cnt = min(base, L.cellCap);
}

}  // namespace

// One CUtensorMap per pyramid level of a buffer set: a 3-D byte tensor {padded columns (pitch), padded rows, frames}
// with strides {pitch, pyrStride}; box = boxW bytes x boxH rows x 1 frame, no swizzle, zero fill outside the tensor.
int orbfe_make_level_maps(const OrbfeFrameGeom& g, const uint8_t* pyr, int frames, int boxW, int boxH, OrbfeFastMaps& maps) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess)
            return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver", cudaGetLastError());
        encode = (EncodeFn)fn;
    }
    memset(&maps, 0, sizeof maps);
    for (int l = 0; l < g.nlevels; l++) {
        const OrbfeLevelGeom& L = g.lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * ORBFE_YOFF), (cuuint64_t)frames};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyrStride};
        const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)std::min(boxH, 256), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(pyr) + L.off, dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for a pyramid level", cudaSuccess);
    }
    return ORBFE_OK;
}

int orbfe_fast_make_maps(const OrbfeFrameGeom& g, OrbfeChunkBufs& b, int frames) {
    return orbfe_make_level_maps(g, b.pyr, frames, (int)RAW_ROW_BYTES, TROWS, b.fastMaps);
}

void orbfe_launch_fast_score(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                             long long* launches) {
    if (g.fastTiles <= 0) return;
    // several tiles per CTA (load of tile i+1 overlapped with the network of tile i) once the grid is large enough
    // to fill the machine several times over; single frames keep one tile per CTA for latency
    const long long tiles = (long long)g.fastTiles * B;
    int per = tiles >= 148LL * 4 * 64 ? 16 : tiles >= 148LL * 4 * 32 ? 8 : tiles >= 148LL * 4 * 16 ? 4 : 1;
    if (const char* ev = getenv("ORBFE_FAST_TILES_PER_CTA")) per = std::max(1, atoi(ev));
    k_fast_score<<<dim3((g.fastTiles + per - 1) / per, B), 256, 0, st>>>(g, b.fastMaps, b.score, per);
    ++*launches;
}

void orbfe_launch_fast_nms(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                           long long* launches) {
    if (g.nmsTiles <= 0) return;
    k_fast_nms<<<dim3(g.nmsTiles, B), 32 * NMS_WARPS, 0, st>>>(g, b.score, b.nmsBits);
    ++*launches;
}

void orbfe_launch_fast_cells(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                             long long* launches) {
    k_fast_cells<<<dim3((g.cellsPerFrame + 7) / 8, B), 256, 0, st>>>(g, b.score, b.nmsBits, b.slots, b.cellCount);
    ++*launches;
}
