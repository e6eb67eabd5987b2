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

constexpr int TW = ORBFE_FAST_TW, TH = ORBFE_FAST_TH, SP = 72;  // smem pitch

__global__ void __launch_bounds__(256)
k_fast_score(const __grid_constant__ OrbfeFrameGeom g, const uint8_t* __restrict__ pyr,
             uint8_t* __restrict__ score) {
    __shared__ __align__(16) uint8_t tile[(TH + 6) * SP];
    int l = 0;
    const int t = blockIdx.x;
    while (l + 1 < g.nlevels && t >= g.lv[l + 1].fastTileBase) l++;
    const OrbfeLevelGeom& L = g.lv[l];
    const int tl = t - L.fastTileBase;
    const int ty = tl / L.fastTilesX, tx = tl - ty * L.fastTilesX;
    // domain origin = ROI (19,19); tile origin in ROI coordinates
    const int ox = 19 + tx * TW, oy = 19 + ty * TH;
    const size_t fo = (size_t)blockIdx.y * g.pyrStride + L.off;
    const uint8_t* src = pyr + fo;
    const int maxx = L.w + 18, maxy = L.h + 18;
    for (int i = threadIdx.x; i < (TH + 6) * (TW + 6); i += 256) {
        const int r = i / (TW + 6), c = i - r * (TW + 6);
        const int x = min(ox - 3 + c, maxx), y = min(oy - 3 + r, maxy);
        tile[r * SP + c] = src[(size_t)(ORBFE_YOFF + y) * L.pitch + ORBFE_XOFF + x];
    }
    __syncthreads();
    const int cx = threadIdx.x & (TW - 1), cy0 = threadIdx.x >> 6;
    const int x = ox + cx;
    if (x >= L.w - 19) return;
    uint8_t* dst = score + fo;
#pragma unroll
    for (int r = 0; r < TH / 4; r++) {
        const int cy = cy0 + 4 * r, y = oy + cy;
        if (y >= L.h - 19) break;
        const uint8_t* p = &tile[(cy + 3) * SP + cx + 3];
        int best = 0;
        if (fc_may_be_corner<SP>(p, g.minTh)) best = fc_arc_best<SP>(p);
        dst[(size_t)(ORBFE_YOFF + y) * L.pitch + ORBFE_XOFF + x] = (uint8_t)(best > g.minTh ? best : 0);
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
    const uint8_t* S = score + (size_t)blockIdx.y * g.pyrStride + L.off + (size_t)ORBFE_YOFF * L.pitch + ORBFE_XOFF;
    uint32_t* out = slots + (size_t)blockIdx.y * g.slotsPerFrame + L.slotBase + (size_t)ci * L.cellCap;
    int nMin = 0, nIni = 0;
    for (int y = y0; y < y1; y++) {
        for (int xb = x0; xb < x1; xb += 32) {
            const int x = xb + lane;
            int c = 0;
            if (x < x1) c = S[(size_t)y * L.pitch + x];
            bool ok = false;
            if (c > 0) {
                int m = 1;  // response must also beat the 0 of empty neighbours: best-1 > 0
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
            const unsigned bi = __ballot_sync(0xffffffffu, ok && c > g.iniTh);
            if (ok) {
                const int pos = nMin + __popc(bm & ((1u << lane) - 1));
                if (pos < L.cellCap)
                    out[pos] = OC_PACK(x - ORBFE_FAST_BORDER, y - ORBFE_FAST_BORDER, c - 1);
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

void orbfe_launch_fast(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                       long long* launches) {
    if (g.fastTiles > 0) {
        k_fast_score<<<dim3(g.fastTiles, B), 256, 0, st>>>(g, b.pyr, b.score);
        ++*launches;
    }
    k_fast_cells<<<dim3((g.cellsPerFrame + 7) / 8, B), 256, 0, st>>>(g, b.score, b.slots, b.cellCount);
    ++*launches;
}
