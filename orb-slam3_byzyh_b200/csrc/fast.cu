// fast.cu -- the FAST half of ORBextractor::ComputeKeyPointsOctTree
// (/root/reference/src/ORBextractor.cc:1061-1165): per 35-px cell cv::FAST(iniThFAST, nms) with
// the minThFAST retry when the cell came back empty, candidates emitted cell row-major and
// FAST row-major inside a cell.
//
// B200 design.  OpenCV's FAST response is threshold independent: with
//   best(p) = max over the 16 arcs of 9 contiguous ring pixels of max(min(c-ring), min(ring-c))
// a pixel is a corner at threshold t iff best > t and its response is best-1.  The 700 tiny
// per-cell cv::FAST calls of the reference therefore collapse into
//   k_fast_score : ONE streaming pass per pyramid level that writes best(p) (0 when
//                  best <= minThFAST) for every pixel of the level's FAST domain, and
//   k_fast_cells : one warp per cell applying the per-cell 3x3 non-max suppression (neighbours
//                  outside the cell's own 3-px-inset interior count as 0, exactly like the
//                  reference's per-cell calls), the "empty at iniThFAST -> retry at minThFAST"
//                  rule and the ordered emission (warp ballot + popc ranks).
// NMS is threshold independent too: a non-corner neighbour has best_n <= t < best, so comparing
// against its true score instead of 0 never changes the outcome; the only threshold-dependent
// step is the final `best > t` filter, which lets one sweep serve both thresholds.
#include "fast_core.h"
#include "octree_core.h"  // OC_PACK
#include "orbfe_internal.h"

namespace {

// ---- k_fast_score -------------------------------------------------------------------------------
// Tile = 128 x 16 output pixels (+3 halo).  The tile is staged in shared memory already widened to
// 16-bit lanes, in four copies shifted by 0..3 pixels, so that ANY run of four horizontally
// adjacent pixels is one aligned LDS.64 (two s16x2 pairs) -- the 16 ring operands of four pixels
// cost 16 LDS.64 and no byte-extraction ALU work, leaving the ALU pipe to the packed min/max
// network of fast_core.h (2 pixels per instruction).  Each thread scores 4 pixels on 2 rows.
constexpr int TW = ORBFE_FAST_TW, TH = ORBFE_FAST_TH;
constexpr int TROWS = TH + 6;         // 22 staged rows
constexpr int TG = TW / 4 + 1;        // 33 four-pixel groups per staged row and copy
static_assert(TW == 128 && TH == 16, "thread mapping below assumes 128x16 tiles");

__global__ void __launch_bounds__(256)
k_fast_score(const __grid_constant__ OrbfeFrameGeom g, const uint8_t* __restrict__ pyr,
             uint8_t* __restrict__ score) {
    __shared__ __align__(16) uint2 cp[4][TROWS][TG];
    int l = 0;
    const int t = blockIdx.x;
    while (l + 1 < g.nlevels && t >= g.lv[l + 1].fastTileBase) l++;
    const OrbfeLevelGeom& L = g.lv[l];
    const int tl = t - L.fastTileBase;
    const int ty = tl / L.fastTilesX, tx = tl - ty * L.fastTilesX;
    // FAST domain origin = ROI (19,19).  Staged column 0 = ROI x 16 + 128*tx = padded column
    // 48 + 128*tx (16-byte aligned); staged row 0 = ROI y 16 + 16*ty.
    const size_t fo = (size_t)blockIdx.y * g.pyrStride + L.off;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pyr + fo);
    const int pw = L.pitch >> 2, w0 = (ORBFE_XOFF + 16 + TW * tx) >> 2;
    const int y0 = ORBFE_YOFF + 16 + TH * ty, ymax = L.h + 2 * ORBFE_YOFF - 1;
    for (int i = threadIdx.x; i < TROWS * TG; i += 256) {
        const int r = i / TG, gq = i - r * TG;
        const uint32_t* row = src + (size_t)min(y0 + r, ymax) * pw;
        const uint32_t a = row[min(w0 + gq, pw - 1)], b = row[min(w0 + gq + 1, pw - 1)];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t v = s ? __funnelshift_r(a, b, 8 * s) : a;   // pixels 4gq+s .. 4gq+s+3
            cp[s][r][gq] = make_uint2(__byte_perm(v, 0u, 0x4140), __byte_perm(v, 0u, 0x4342));
        }
    }
    __syncthreads();
    const int gq = threadIdx.x & 31, rp = threadIdx.x >> 5;
    const int x = 19 + TW * tx + 4 * gq;       // ROI x of the first of this thread's 4 pixels
    if (x >= L.w - 19) return;
    const uint32_t sub2 = (uint32_t)g.minTh * 0x00010001u;
    // score map column = ROI x + 13, so that a 4-pixel group is one aligned 32-bit store
    uint8_t* dst = score + fo + ORBFE_SXOFF + x;
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int orow = 2 * rp + rr, y = 19 + TH * ty + orow;
        if (y >= L.h - 19) break;
        const uint2 c = cp[3][orow + 3][gq];
        const uint32_t c0 = c.x + FC_BIAS2, c1 = c.y + FC_BIAS2;
        uint32_t e0[16], e1[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int o = 3 + FC_RING_DX(k);
            const uint2 v = cp[o & 3][orow + 3 + FC_RING_DY(k)][gq + (o >> 2)];
            e0[k] = c0 - v.x;   // both lanes stay in [1, 511]: no borrow crosses the lane boundary
            e1[k] = c1 - v.y;
        }
        const uint32_t m0 = fc_margin2(e0, sub2), m1 = fc_margin2(e1, sub2);
        *reinterpret_cast<uint32_t*>(dst + (size_t)(ORBFE_YOFF + y) * L.pitch) = __byte_perm(m0, m1, 0x6420);
    }
}

__global__ void __launch_bounds__(256)
k_fast_cells(const __grid_constant__ OrbfeFrameGeom g, const uint8_t* __restrict__ score,
             uint32_t* __restrict__ slots, int* __restrict__ cellCount) {
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
    const int x0 = iniX + 3, x1 = maxX - 3, y0 = iniY + 3, y1 = maxY - 3;  // FAST interior
    // score map: margin = best - minThFAST (0 = no corner), column = ROI x + ORBFE_SXOFF
    const uint8_t* S = score + (size_t)blockIdx.y * g.pyrStride + L.off + (size_t)ORBFE_YOFF * L.pitch + ORBFE_SXOFF;
    const int iniMargin = g.iniTh - g.minTh;   // best > iniThFAST  <=>  margin > iniTh - minTh
    uint32_t* out = slots + (size_t)blockIdx.y * g.slotsPerFrame + L.slotBase + (size_t)ci * L.cellCap;
    int nMin = 0, nIni = 0;
    for (int y = y0; y < y1; y++) {
        for (int xb = x0; xb < x1; xb += 32) {
            const int x = xb + lane;
            int c = 0;
            if (x < x1) c = S[(size_t)y * L.pitch + x];
            bool ok = false;
            if (c > 0) {
                // response (best-1) must also beat the 0 of empty neighbours: best > 1
                int m = max(1 - g.minTh, 0);
#pragma unroll
                for (int dy = -1; dy <= 1; dy++) {
                    const int yy = y + dy;
                    if (yy < y0 || yy >= y1) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++) {
                        const int xx = x + dx;
                        if ((dx | dy) == 0 || xx < x0 || xx >= x1) continue;
                        m = max(m, (int)S[(size_t)yy * L.pitch + xx]);
                    }
                }
                ok = c > m;
            }
            const unsigned bm = __ballot_sync(0xffffffffu, ok);
            const unsigned bi = __ballot_sync(0xffffffffu, ok && c > iniMargin);
            if (ok) {
                const int pos = nMin + __popc(bm & ((1u << lane) - 1));
                if (pos < L.cellCap)
                    out[pos] = OC_PACK(x - ORBFE_FAST_BORDER, y - ORBFE_FAST_BORDER, c + g.minTh - 1);
            }
            nMin += __popc(bm);
            nIni += __popc(bi);
        }
    }
    nMin = min(nMin, L.cellCap);
    int count = nMin;
    if (nIni > 0 && nIni < nMin) {
        // the cell was not empty at iniThFAST: keep only those corners (ordered compaction)
        __syncwarp();
        int wpos = 0;
        for (int base = 0; base < nMin; base += 32) {
            const int k = base + lane;
            uint32_t v = 0;
            bool keep = false;
            if (k < nMin) {
                v = out[k];
                keep = OC_PK_S(v) + 1 > g.iniTh;
            }
            const unsigned bk = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) out[wpos + __popc(bk & ((1u << lane) - 1))] = v;
            wpos += __popc(bk);
            __syncwarp();
        }
        count = wpos;
    }
    if (lane == 0) *cnt = count;
}

}  // namespace

void orbfe_launch_fast_score(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                             long long* launches) {
    if (g.fastTiles <= 0) return;
    k_fast_score<<<dim3(g.fastTiles, B), 256, 0, st>>>(g, b.pyr, b.score);
    ++*launches;
}

void orbfe_launch_fast_cells(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                             long long* launches) {
    k_fast_cells<<<dim3((g.cellsPerFrame + 7) / 8, B), 256, 0, st>>>(g, b.score, b.slots, b.cellCount);
    ++*launches;
}
