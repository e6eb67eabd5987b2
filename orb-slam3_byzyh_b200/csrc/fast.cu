// fast.cu -- the FAST half of ORBextractor::ComputeKeyPointsOctTree
// (/root/reference/src/ORBextractor.cc:1061-1165): per 35-px cell cv::FAST(iniThFAST, nms) with
// the minThFAST retry when the cell came back empty, candidates emitted cell row-major and
// FAST row-major inside a cell.
//
// B200 design: ONE kernel, one WARP per cell, no score map in HBM.  The reference's cell is the natural unit:
// the cell interiors tile the FAST domain exactly, a neighbour outside the cell's own interior counts as 0 in
// cv::FAST's 3x3 non-max suppression, and the minThFAST retry is a per-cell decision -- so a warp that owns a
// cell (interior <= 69 x 69 px, + the 3-px ring halo = one TMA box of the padded level) can run the reference's
// control flow literally: FAST at iniThFAST, NMS, "empty?", FAST at minThFAST, ordered emission.  The level is
// read from HBM once (cp.async.bulk.tensor, one box per cell), nothing but the candidate slots is written.
//
// Per cell and threshold t (k_fast_cells):
//   1. dense early reject, 8 px per lane on packed bytes (fast_core.h: fc_compass4).  Every arc of 9 ring pixels
//      contains a pixel of each opposite pair {k, k+8}; the two compass pairs are tested with VABSDIFF4 + a SWAR
//      byte compare.  On the bench frames 36 % of the pixels survive at t = 20 (13 % are corners), on camera
//      images a few per cent.
//   2. survivors are compacted into a shared-memory queue (warp scan of the per-lane counts); whenever 64 are
//      queued they are scored two per lane on the u16x2 min/max network of fast_core.h (arcs taken in pairs on the
//      raw ring values), so the network always runs on full lanes.  Scores > t go into a byte tile S (zero
//      elsewhere, zero border) and their positions into a corner list.
//   3. NMS per listed corner against its 8 neighbours in S (strictly greater, cv::FAST's rule); kept corners set
//      a bit in a per-row bitmap of the cell.
//   4. the cell is empty at iniThFAST -> the same again at minThFAST (:1141-1148); then lane = row, popcounts +
//      warp scan give every row its output rank, the bits are emitted row-major (= cv::FAST's order) into the
//      cell's candidate slots.
// The response is threshold independent (corner at t <=> best > t, response = best - 1), and so is the NMS: a
// non-corner neighbour has best <= t < best of the corner, so leaving it at 0 never changes an outcome.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fast_core.h"
#include "tma.h"
#include "octree_core.h"  // OC_PACK
#include "orbfe_internal.h"

namespace {

constexpr int FW = 8;                 // warps (= cells in flight) per CTA
constexpr int ROW0 = 3;               // staged row of the cell's first interior row (3-px ring halo above it)
// Staged column of the first interior pixel: `cs` in [3, 18], per cell.  The TMA unit only takes box origins whose
// innermost coordinate is a multiple of 16 bytes (measured on B200: any other origin raises "illegal instruction"),
// so the box starts at the 16-byte boundary at or below the left ring pixel.
constexpr int CSMAX = 18;
constexpr int QROUND = 64;            // survivors scored per round (two per lane)
constexpr int QCAP = 704;             // survivor queue (a whole 36 x 38 cell of the bench frames fits: one flush per cell)
constexpr int QFLUSH = QCAP - 256;    // flush before a dense step (8 px x 32 lanes) could overflow it
constexpr int CCAP = 512;             // corner list (more corners than that: the NMS walks the whole interior instead)

struct FastLayout {   // per-warp shared-memory carve-up, derived on the host from the geometry
    int rawBytes;     // TMA box of the largest cell (multiple of 128)
    int sBytes;       // score tile: the same geometry without the two outermost rows at either end
    int stride;       // bytes per warp
};

// TMA box of one cell of level L: pitch x rows of the staged tile.
__host__ __device__ inline int fast_box_w(const OrbfeLevelGeom& L) { return (CSMAX + L.wCell + 3 + 15) & ~15; }
__host__ __device__ inline int fast_box_h(const OrbfeLevelGeom& L) { return L.hCell + 6; }

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)   // SHFL.UP with predicate out + predicated add: two instructions per step
        asm volatile("{\n\t.reg .pred p;\n\t.reg .s32 t;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.s32 %0, %0, t;\n\t}"
                     : "+r"(v) : "r"(o));
    return v;
}

// Scores up to 64 queued survivors Q[base .. base+n): lane takes entries base+lane and base+32+lane, one per 16-bit
// half of the packed network.  An entry is the byte offset of the pixel inside the staged tile; Sb is the score tile
// addressed by the same offsets.  Corners (best > t) are appended to C in queue order.
template <int P>
__device__ __forceinline__ void fast_score_round(const uint8_t* __restrict__ raw, uint8_t* __restrict__ Sb,
                                                 const uint16_t* __restrict__ Q, int base, int n, int t,
                                                 uint16_t* __restrict__ C, int& nC, int lane) {
    const bool v0 = lane < n, v1 = lane + 32 < n;
    const int dummy = ROW0 * P + 8;   // a pixel whose ring lies inside the tile
    const int e0 = v0 ? (int)Q[base + lane] : dummy, e1 = v1 ? (int)Q[base + 32 + lane] : dummy;
    const uint8_t* p0 = raw + e0;
    const uint8_t* p1 = raw + e1;
    uint32_t r[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int o = FC_RING_DX(k) + FC_RING_DY(k) * P;
        r[k] = (uint32_t)p0[o] + ((uint32_t)p1[o] << 16);   // multiply-add: keeps the pack off the ALU pipe of the network
    }
    const uint32_t c2 = (uint32_t)p0[0] + ((uint32_t)p1[0] << 16);
    const uint32_t m = fc_margin2_pair_raw_biased(r, c2, (uint32_t)t * 0x00010001u);   // per half: max(best - t, 0)
    const int m0 = (int)(m & 0xFFFFu), m1 = (int)(m >> 16);
    const bool k0 = v0 && m0 > 0, k1 = v1 && m1 > 0;
    if (k0) Sb[e0] = (uint8_t)(m0 + t);
    if (k1) Sb[e1] = (uint8_t)(m1 + t);
    // queue order = entries base .. base+31 (first halves) then base+32 .. base+63 (second halves)
    const unsigned b0 = __ballot_sync(0xffffffffu, k0), b1 = __ballot_sync(0xffffffffu, k1);
    const unsigned lt = (1u << lane) - 1u;
    const int q0 = nC + __popc(b0 & lt), q1 = nC + __popc(b0) + __popc(b1 & lt);
    if (k0 && q0 < CCAP) C[q0] = (uint16_t)e0;
    if (k1 && q1 < CCAP) C[q1] = (uint16_t)e1;
    nC += __popc(b0) + __popc(b1);
}

// Scores the queue front to back in rounds of 64.  Not final: the < 64 entries left over move to the front (they
// keep their order) and their number is returned; final: they get a last, partly filled round.
template <int P>
__device__ __forceinline__ int fast_flush(const uint8_t* __restrict__ raw, uint8_t* __restrict__ Sb, uint16_t* __restrict__ Q,
                                          int tail, bool final, int t, uint16_t* __restrict__ C, int& nC, int lane) {
    int base = 0;
    for (; base + QROUND <= tail; base += QROUND) fast_score_round<P>(raw, Sb, Q, base, QROUND, t, C, nC, lane);
    const int left = tail - base;
    if (final) {
        if (left > 0) fast_score_round<P>(raw, Sb, Q, base, left, t, C, nC, lane);
        return 0;
    }
    if (base > 0 && left > 0) {
        const uint16_t a = lane < left ? Q[base + lane] : (uint16_t)0, b = lane + 32 < left ? Q[base + 32 + lane] : (uint16_t)0;
        __syncwarp();
        if (lane < left) Q[lane] = a;
        if (lane + 32 < left) Q[32 + lane] = b;
    }
    return left;
}

// One corner against its eight neighbours in the score tile (cv::FAST keeps it iff it is strictly greater).
template <int P>
__device__ __forceinline__ bool fast_is_max(const uint8_t* __restrict__ s, int v) {
    const int a = max(max((int)s[-1], (int)s[1]), max((int)s[-P], (int)s[P]));
    const int b = max(max((int)s[-P - 1], (int)s[-P + 1]), max((int)s[P - 1], (int)s[P + 1]));
    return v > max(a, b);
}

// cv::FAST(cell, t, nms = true): dense reject -> queue -> score -> corner list -> NMS + ordered emission into `out`.
// The score tile is zero on entry.  Returns the number of keypoints.
template <int P, bool SMALLU>
__device__ __forceinline__ int fast_cell_pass(const uint8_t* __restrict__ raw, uint8_t* __restrict__ Sb,
                                              uint16_t* __restrict__ Q, uint16_t* __restrict__ C, int cs, int nbx, int nby,
                                              int t, int x0, int y0, uint32_t* __restrict__ out, int cellCap, int lane) {
    if (t >= 255) return 0;   // best <= 255: no pixel is a corner
    const int u = t + 1;      // |ring - centre| >= u
    const uint32_t uLow = (uint32_t)(u & 0x7F) * 0x01010101u, uTop = (u & 0x80) ? 0xFFFFFFFFu : 0u;
    // A lane takes an aligned octet of staged columns (one LDS.64); the interior is the columns [cs, cs + nbx).  Lane
    // = (row of the step, octet of the row): the lane's column, and with it the masks of the pixels outside the
    // interior, stay fixed while the warp walks down the cell `rps` rows per step; queue order is row-major.
    const int a0 = cs & ~7, lo0 = cs - a0;
    const int opr = (lo0 + nbx + 7) >> 3;                // octets per interior row (<= 11)
    const int rps = 32 / opr;                            // rows per step
    const int r = lane / opr, oc = lane - r * opr;
    const bool active = r < rps;
    const int rem = lo0 + nbx - 8 * oc, lo = max(lo0 - 8 * oc, 0);
    const uint32_t mk0 = active ? __funnelshift_rc(0x80808080u, 0u, 8 * max(4 - rem, 0)) & __funnelshift_lc(0u, 0x80808080u, 8 * min(lo, 4)) : 0u;
    const uint32_t mk1 = active ? __funnelshift_rc(0x80808080u, 0u, 8 * max(8 - rem, 0)) & __funnelshift_lc(0u, 0x80808080u, 8 * max(lo - 4, 0)) : 0u;
    int e = ((active ? r : 0) + ROW0) * P + a0 + 8 * oc;  // tile offset of the octet's first pixel (8-byte aligned)
    int tail = 0, nC = 0;
    for (int row0 = 0; row0 < nby; row0 += rps, e += rps * P) {
        // rows past the cell's last one read the rows below the tile (the score tile follows it): masked
        const bool rowok = row0 + r < nby;
        const uint2 c = *reinterpret_cast<const uint2*>(raw + e);
        const uint2 no = *reinterpret_cast<const uint2*>(raw + e - 3 * P);
        const uint2 so = *reinterpret_cast<const uint2*>(raw + e + 3 * P);
        const uint32_t wm = *reinterpret_cast<const uint32_t*>(raw + e - 4);
        const uint32_t wp = *reinterpret_cast<const uint32_t*>(raw + e + 8);
        uint32_t f0 = fc_compass4<SMALLU>(c.x, no.x, so.x, __funnelshift_r(wm, c.x, 8), __funnelshift_r(c.x, c.y, 24), uLow, uTop);
        uint32_t f1 = fc_compass4<SMALLU>(c.y, no.y, so.y, __funnelshift_r(c.x, c.y, 8), __funnelshift_r(c.y, wp, 24), uLow, uTop);
        f0 &= rowok ? mk0 : 0u;
        f1 &= rowok ? mk1 : 0u;
        const int cnt = __popc(f0 | (f1 >> 1));          // bits 7 and 6 of every byte
        const int incl = warp_incl_scan(cnt);
        uint16_t* q = Q + tail + incl - cnt;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (f0 & (0x80u << (8 * j))) *q++ = (uint16_t)(e + j);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (f1 & (0x80u << (8 * j))) *q++ = (uint16_t)(e + 4 + j);
        tail += __shfl_sync(0xffffffffu, incl, 31);
        if (tail > QFLUSH) {
            __syncwarp();
            tail = fast_flush<P>(raw, Sb, Q, tail, false, t, C, nC, lane);
            __syncwarp();
        }
    }
    __syncwarp();
    fast_flush<P>(raw, Sb, Q, tail, true, t, C, nC, lane);
    __syncwarp();
    // ---- 3x3 non-max suppression + emission.  The score tile is 0 outside the cell interior, which is cv::FAST's
    // rule for a cell-sized image; C is in row-major order (the queue's), so a ballot ranks the kept corners in
    // cv::FAST's emission order. ----
    int count = 0;
    const unsigned lt = (1u << lane) - 1u;
    if (nC <= CCAP) {
        for (int i0 = 0; i0 < nC; i0 += 32) {
            const bool ok = i0 + lane < nC;
            const int ec = ok ? (int)C[i0 + lane] : ROW0 * P + 8;
            const int v = Sb[ec];
            const bool keep = ok && fast_is_max<P>(Sb + ec, v);
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            const int pos = count + __popc(b & lt);
            if (keep && pos < cellCap) {
                const int row = ec / P, col = ec - row * P - cs;
                // window coordinates (origin = minBorder 16): ROI - 16; response = best - 1
                out[pos] = OC_PACK(x0 + col - ORBFE_FAST_BORDER, y0 + row - ROW0 - ORBFE_FAST_BORDER, v - 1);
            }
            count += __popc(b);
        }
    } else {   // more corners than the list holds (noise-like cells): walk the whole interior, row-major
        const int npx = nbx * nby;
        for (int i0 = 0; i0 < npx; i0 += 32) {
            const int i = min(i0 + lane, npx - 1);
            const int row = i / nbx, col = i - row * nbx;
            const uint8_t* s = Sb + (row + ROW0) * P + cs + col;
            const int v = s[0];
            const bool keep = i0 + lane < npx && v > 0 && fast_is_max<P>(s, v);
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            const int pos = count + __popc(b & lt);
            if (keep && pos < cellCap) out[pos] = OC_PACK(x0 + col - ORBFE_FAST_BORDER, y0 + row - ORBFE_FAST_BORDER, v - 1);
            count += __popc(b);
        }
    }
    return count;
}

__device__ __forceinline__ void fast_zero(uint8_t* S, int bytes16, int lane) {
    uint4* s4 = reinterpret_cast<uint4*>(S);
    for (int i = lane; i < bytes16; i += 32) s4[i] = make_uint4(0u, 0u, 0u, 0u);
}

// The reference's per-cell body (:1135-1165) on a staged cell.  Returns the number of candidates.
template <int P>
__device__ __forceinline__ int fast_cell(const uint8_t* __restrict__ raw, uint8_t* __restrict__ S, uint16_t* __restrict__ Q,
                                         uint16_t* __restrict__ C, int boxH, int cs, int nbx, int nby, int x0, int y0,
                                         int iniTh, int minTh, uint32_t* __restrict__ out, int cellCap, int lane) {
    // cv::FAST's response is best - 1 and its NMS compares responses, so a corner with best == 1 (possible only at
    // threshold 0) scores 0 like a non-corner: it never wins and never blocks.  Threshold 0 therefore acts as 1.
    iniTh = max(iniTh, 1);
    minTh = max(minTh, 1);
    uint8_t* Sb = S - 2 * P;   // the score tile starts at staged row 2: same offsets as the staged tile
    // thresholds below 127 (all the reference ever uses) take the cheaper byte compare of the dense reject
    int total = iniTh < 127 ? fast_cell_pass<P, true>(raw, Sb, Q, C, cs, nbx, nby, iniTh, x0, y0, out, cellCap, lane)
                            : fast_cell_pass<P, false>(raw, Sb, Q, C, cs, nbx, nby, iniTh, x0, y0, out, cellCap, lane);
    // vKeysCell.empty() -> FAST(minThFAST) (:1141-1148).  With minThFAST >= iniThFAST the retry cannot find anything:
    // its corners are a subset and the NMS outcome of a corner does not depend on the threshold.
    if (total == 0 && minTh < iniTh) {
        __syncwarp();
        fast_zero(S, P * (boxH - 4) / 16, lane);
        __syncwarp();
        total = minTh < 127 ? fast_cell_pass<P, true>(raw, Sb, Q, C, cs, nbx, nby, minTh, x0, y0, out, cellCap, lane)
                            : fast_cell_pass<P, false>(raw, Sb, Q, C, cs, nbx, nby, minTh, x0, y0, out, cellCap, lane);
    }
    return total;
}

__global__ void __launch_bounds__(32 * FW, 3)
k_fast_cells(const __grid_constant__ OrbfeFrameGeom g, const __grid_constant__ OrbfeFastMaps maps,
             const OrbfeFastCell* __restrict__ cells, uint32_t* __restrict__ slots, int* __restrict__ cellCount,
             const FastLayout lay, int cellsPerWarp) {
    extern __shared__ __align__(128) uint8_t fsm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* raw = fsm + (size_t)wid * lay.stride;
    uint8_t* S = raw + lay.rawBytes;
    uint16_t* Q = reinterpret_cast<uint16_t*>(S + lay.sBytes);
    uint16_t* C = Q + QCAP;
    uint64_t* bar = reinterpret_cast<uint64_t*>(C + CCAP);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncwarp();
    const int frame = blockIdx.y;
    const int c0 = (blockIdx.x * FW + wid) * cellsPerWarp, c1 = min(c0 + cellsPerWarp, g.cellsPerFrame);
    uint32_t parity = 0;
    for (int cell = c0; cell < c1; cell++) {
        // cell geometry (ORBextractor.cc:1098-1132), derived once per geometry on the host (orbfe_fast_cell_table)
        const OrbfeFastCell fc = cells[cell];
        int* cnt = cellCount + (size_t)frame * g.cellsPerFrame + cell;
        if (fc.nbx == 0) {   // the reference skips the cell, or cv::FAST has no interior pixel in it
            if (lane == 0) *cnt = 0;
            continue;
        }
        const int l = fc.level;
        const OrbfeLevelGeom& L = g.lv[l];
        const int ci = cell - L.cellBase;
        const int x0 = fc.x0, y0 = fc.y0, nbx = fc.nbx, nby = fc.nby;   // FAST interior (ROI coordinates)
        const int bw = fast_box_w(L), bh = fast_box_h(L);
        const int xs = (ORBFE_XOFF + x0 - 3) & ~15, cs = ORBFE_XOFF + x0 - xs;   // box origin (padded column), staged column of x0
        if (lane == 0) {
            // the generic-proxy reads of the previous cell (ordered by the __syncwarp at the loop end) come before
            // the async-proxy writes of this request
            fence_proxy_async();
            mbar_expect_tx(bar, (uint32_t)(bw * bh));
            tma_tile_g2s(raw, &maps.m[l], xs, ORBFE_YOFF + y0 - ROW0, frame, bar);
        }
        fast_zero(S, bw * (bh - 4) / 16, lane);   // under the copy
        mbar_wait(bar, parity);
        parity ^= 1u;
        __syncwarp();
        uint32_t* out = slots + (size_t)frame * g.slotsPerFrame + L.slotBase + (size_t)ci * L.cellCap;
        int total;
        if (bw == 64)
            total = fast_cell<64>(raw, S, Q, C, bh, cs, nbx, nby, x0, y0, g.iniTh, g.minTh, out, L.cellCap, lane);
        else if (bw == 80)
            total = fast_cell<80>(raw, S, Q, C, bh, cs, nbx, nby, x0, y0, g.iniTh, g.minTh, out, L.cellCap, lane);
        else
            total = fast_cell<96>(raw, S, Q, C, bh, cs, nbx, nby, x0, y0, g.iniTh, g.minTh, out, L.cellCap, lane);
        if (lane == 0) *cnt = min(total, L.cellCap);
        __syncwarp();
    }
}

FastLayout fast_layout(const OrbfeFrameGeom& g) {
    FastLayout lay = {};
    int raw = 128, sb = 128;
    for (int l = 0; l < g.nlevels; l++) {
        const OrbfeLevelGeom& L = g.lv[l];
        if (L.nCols == 0) continue;
        raw = std::max(raw, fast_box_w(L) * fast_box_h(L));
        sb = std::max(sb, fast_box_w(L) * (fast_box_h(L) - 4));
    }
    lay.rawBytes = (raw + 127) & ~127;
    lay.sBytes = (sb + 15) & ~15;
    lay.stride = (lay.rawBytes + lay.sBytes + 2 * QCAP + 2 * CCAP + 16 + 127) & ~127;
    return lay;
}

}  // namespace

// One CUtensorMap per pyramid level of a buffer set: a 3-D byte tensor {padded columns (pitch), padded rows, frames}
// with strides {pitch, pyrStride}; box = boxW[l] bytes x boxH[l] rows x 1 frame, no swizzle, zero fill outside the tensor.
int orbfe_make_level_maps(const OrbfeFrameGeom& g, const uint8_t* pyr, int frames, const int* boxW, const int* boxH,
                          OrbfeFastMaps& maps) {
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
        const cuuint32_t box[3] = {(cuuint32_t)std::min(std::max(boxW[l], 16), 256), (cuuint32_t)std::min(std::max(boxH[l], 1), 256), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(pyr) + L.off, dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for a pyramid level", cudaSuccess);
    }
    return ORBFE_OK;
}

int orbfe_fast_make_maps(const OrbfeFrameGeom& g, OrbfeChunkBufs& b, int frames) {
    int bw[ORBFE_MAX_LEVELS], bh[ORBFE_MAX_LEVELS];
    for (int l = 0; l < g.nlevels; l++) {
        bw[l] = g.lv[l].nCols ? fast_box_w(g.lv[l]) : 16;
        bh[l] = g.lv[l].nCols ? fast_box_h(g.lv[l]) : 1;
        if (bw[l] > 96 || bh[l] > 256) return orbfe_fail(ORBFE_ERR_INVALID, "FAST cell larger than the staged tile", cudaSuccess);
    }
    const FastLayout lay = fast_layout(g);
    if (cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, FW * lay.stride) != cudaSuccess)
        return orbfe_fail(ORBFE_ERR_CUDA, "cudaFuncSetAttribute(k_fast_cells)", cudaGetLastError());
    return orbfe_make_level_maps(g, b.pyr, frames, bw, bh, b.fastMaps);
}

// The reference's cell loop (:1098-1132) on the host: per cell the FAST interior [x0, x0 + nbx) x [y0, y0 + nby) in ROI
// coordinates, nbx == 0 where the reference skips the cell (:1108, :1125) or the cell image is narrower than FAST's 7 px.
void orbfe_fast_cell_table(const OrbfeFrameGeom& g, std::vector<OrbfeFastCell>& out) {
    out.assign((size_t)std::max(g.cellsPerFrame, 1), OrbfeFastCell{0, 0, 0, 0, 0, 0});
    for (int l = 0; l < g.nlevels; l++) {
        const OrbfeLevelGeom& L = g.lv[l];
        for (int i = 0; i < L.nRows; i++)
            for (int j = 0; j < L.nCols; j++) {
                OrbfeFastCell& c = out[(size_t)L.cellBase + (size_t)i * L.nCols + j];
                c.level = (uint8_t)l;
                const int iniY = ORBFE_FAST_BORDER + i * L.hCell, iniX = ORBFE_FAST_BORDER + j * L.wCell;
                const int maxY = std::min(iniY + L.hCell + 6, L.maxBY), maxX = std::min(iniX + L.wCell + 6, L.maxBX);
                if (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6 || maxX - iniX < 7 || maxY - iniY < 7) continue;
                c.x0 = (uint16_t)(iniX + 3);
                c.y0 = (uint16_t)(iniY + 3);
                c.nbx = (uint8_t)(maxX - iniX - 6);
                c.nby = (uint8_t)(maxY - iniY - 6);
            }
    }
}

void orbfe_launch_fast(const OrbfeFrameGeom& g, const OrbfeFastCell* cells, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                       long long* launches) {
    if (g.cellsPerFrame <= 0) return;
    const FastLayout lay = fast_layout(g);
    // several cells per warp once the grid fills the machine many times over; single frames keep one cell per warp
    const long long ncell = (long long)g.cellsPerFrame * B;
    int per = ncell >= 148LL * 28 * 64 ? 8 : ncell >= 148LL * 28 * 16 ? 4 : ncell >= 148LL * 28 * 4 ? 2 : 1;
    if (const char* ev = getenv("ORBFE_FAST_CELLS_PER_WARP")) per = std::max(1, atoi(ev));
    const int gx = (g.cellsPerFrame + FW * per - 1) / (FW * per);
    k_fast_cells<<<dim3(gx, B), 32 * FW, FW * lay.stride, st>>>(g, b.fastMaps, cells, b.slots, b.cellCount, lay, per);
    ++*launches;
}
