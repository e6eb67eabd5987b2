// orbfe_api.cu -- host planner and C ABI of the extraction path (include/orbfe.h).
//
// Mirrors ORB_SLAM3::ORBextractor (/root/reference/include/ORBextractor.h:46-112,
// src/ORBextractor.cc:468-571 ctor, :1557-1682 operator()) for batches of frames: every table the
// reference derives in its constructor or per call (scale pyramid, features per level, level
// sizes, FAST cell grid, quadtree roots) is derived here on the host with the same arithmetic
// and handed to the kernels as one by-value geometry block; the per-pixel / per-keypoint work is
// CUDA only.  There is no CPU fallback: without a usable device every call returns ORBFE_ERR_CUDA.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "orbfe_internal.h"
#include "scratch.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return fail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

int orbfe_fail(int code, const char* what, cudaError_t e) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}

namespace {

// reference src/ORBextractor.cc:468-527
void build_tables(OrbfeExtractor* e) {
    const int n = e->nlevels;
    e->scale.assign(n, 1.f);
    e->sigma2.assign(n, 1.f);
    for (int i = 1; i < n; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    e->invScale.resize(n);
    e->invSigma2.resize(n);
    for (int i = 0; i < n; i++) {
        e->invScale[i] = 1.0f / e->scale[i];
        e->invSigma2[i] = 1.0f / e->sigma2[i];
    }
    e->nfeat.assign(n, 0);
    const float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = (float)(e->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n)));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) {
        e->nfeat[l] = cv_round(nDesired);
        sum += e->nfeat[l];
        nDesired *= factor;
    }
    e->nfeat[n - 1] = std::max(e->nfeatures - sum, 0);
}

// OpenCV resize(INTER_LINEAR, CV_8U) coefficient tables (see oracle/cvprims.cpp for the pinned
// restatement): index/weights per destination column (clamped both ends) and row.
void build_taps(int ssize, int dsize, bool isX, OrbfeTap* out) {
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        int s0, s1;
        if (isX) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
            s0 = s;
            s1 = std::min(s + 1, ssize - 1);
        } else {
            s0 = std::min(std::max(s, 0), ssize - 1);
            s1 = std::min(std::max(s + 1, 0), ssize - 1);
        }
        out[d].s = (short)s0;
        out[d].s1 = (short)s1;
        out[d].a0 = (short)cv_round((1.f - f) * 2048.f);
        out[d].a1 = (short)cv_round(f * 2048.f);
    }
}

// Every stream of the handle idle: before anything the batches in flight may still touch is freed or re-laid out.
int quiesce(OrbfeExtractor* e) {
    if (e->sH2D) CK(cudaStreamSynchronize(e->sH2D));
    CK(cudaStreamSynchronize(e->sCompute));
    if (e->sCompute2) CK(cudaStreamSynchronize(e->sCompute2));
    if (e->sD2H) CK(cudaStreamSynchronize(e->sD2H));
    return ORBFE_OK;
}

// Root nodes of DistributeOctTree on level l of a rows x cols frame (:718: round(width / height) of the level's FAST
// window); 1 for a level without a full 35-px cell (no keypoints there), < 1 where the reference divides by zero.
int level_roots(const OrbfeExtractor* e, int rows, int cols, int l) {
    const int w = cv_round((float)cols * e->invScale[l]), h = cv_round((float)rows * e->invScale[l]);
    const float width = (float)(w - 2 * ORBFE_FAST_BORDER), height = (float)(h - 2 * ORBFE_FAST_BORDER);
    if (!(width >= 35.f && height >= 35.f)) return 1;
    return (int)roundf(width / height);
}

int build_geometry(OrbfeExtractor* e, int rows, int cols) {
    if (e->haveGeom && e->g.rows == rows && e->g.cols == cols) return ORBFE_OK;
    if (rows > 4096 || cols > 4096) return fail(ORBFE_ERR_INVALID, "image larger than 4096 px");
    OrbfeFrameGeom g;
    memset(&g, 0, sizeof g);
    g.nlevels = e->nlevels;
    g.rows = rows;
    g.cols = cols;
    g.iniTh = std::min(std::max(e->iniTh, 0), 255);  // cv::FAST clamps its threshold
    g.minTh = std::min(std::max(e->minTh, 0), 255);
    size_t off = 0;
    unsigned slot = 0, tapOff = 0;
    int cell = 0, kpBase = 0, blurTile = 0;
    std::vector<OrbfeTap> taps;
    for (int l = 0; l < g.nlevels; l++) {
        OrbfeLevelGeom& L = g.lv[l];
        const float sc = e->invScale[l];
        L.w = cv_round((float)cols * sc);  // :1692
        L.h = cv_round((float)rows * sc);
        L.maxBX = L.w - ORBFE_FAST_BORDER;  // :1076-1079
        L.maxBY = L.h - ORBFE_FAST_BORDER;
        if (L.w < 1 || L.h < 1) return fail(ORBFE_ERR_INVALID, "image too small for nlevels (a pyramid level is empty)");
        const float width = (float)(L.maxBX - ORBFE_FAST_BORDER), height = (float)(L.maxBY - ORBFE_FAST_BORDER);
        if (width <= 0.f || height <= 0.f)
            return fail(ORBFE_ERR_INVALID, "a pyramid level is smaller than the 16-px FAST border (the reference aborts there)");
        // A level narrower than one 35-px cell has nCols or nRows == 0 in the reference: its cell
        // loops do not run and the level simply yields no keypoints (:1087-1098).
        const bool hasCells = width >= 35.f && height >= 35.f;
        L.nCols = hasCells ? (int)(width / 35.f) : 0;  // :1087-1095
        L.nRows = hasCells ? (int)(height / 35.f) : 0;
        L.wCell = hasCells ? (int)ceilf(width / L.nCols) : 0;
        L.hCell = hasCells ? (int)ceilf(height / L.nRows) : 0;
        L.pitch = (int)align_up(ORBFE_XOFF + L.w + ORBFE_EDGE, 16);
        L.off = (unsigned)off;
        off += align_up((size_t)L.pitch * (L.h + 2 * ORBFE_YOFF), 256);
        L.cellBase = cell;
        cell += L.nCols * L.nRows;
        L.cellCap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
        L.slotBase = slot;
        L.candCap = L.nCols * L.nRows * L.cellCap;
        slot += (unsigned)L.candCap;
        L.nfeat = e->nfeat[l];
        L.nIni = level_roots(e, rows, cols, l);  // :718
        if (L.nIni < 1) return fail(ORBFE_ERR_INVALID, "aspect ratio < 0.5 (the reference divides by zero)");
        L.hX = hasCells ? width / L.nIni : 1.f;
        L.ocM = std::max(L.nfeat + 3, 4 * L.nIni) + 1;
        L.kpBase = kpBase;
        L.kpCap = L.ocM;
        kpBase += L.kpCap;
        L.scale = e->scale[l];
        L.invScale = e->invScale[l];
        L.kpsize = (float)(int)(31 * e->scale[l]);  // :1184
        L.mode = 0;
        if (l > 0) {
            const OrbfeLevelGeom& S = g.lv[l - 1];
            const double sx = 1.0 / ((double)L.w / S.w), sy = 1.0 / ((double)L.h / S.h);
            if (L.w == S.w && L.h == S.h) L.mode = 2;
            else if (fabs(sx - 2.0) < 2.220446049250313e-16 && fabs(sy - 2.0) < 2.220446049250313e-16) L.mode = 1;
            L.xtab = tapOff;
            L.ytab = tapOff + (unsigned)L.w;
            taps.resize(tapOff + L.w + L.h);
            build_taps(S.w, L.w, true, taps.data() + L.xtab);
            build_taps(S.h, L.h, false, taps.data() + L.ytab);
            tapOff += (unsigned)(L.w + L.h);
            // k_resize_fast precondition, checked on the padded (reflected) destination columns
            L.fastTaps = 1;
            for (int wc = 0; wc < L.pitch / 4 && L.fastTaps; wc++) {
                int lo = 1 << 30, hi = 0;
                for (int i = 0; i < 4; i++) {
                    int p = 4 * wc - ORBFE_XOFF + i;
                    if (p < 0) p = -p;
                    if (p >= L.w) p = 2 * (L.w - 1) - p;
                    p = std::max(0, std::min(p, L.w - 1));
                    const OrbfeTap& t = taps[L.xtab + p];
                    lo = std::min(lo, (int)t.s);
                    hi = std::max(hi, (int)t.s1);
                }
                if (hi - lo > 7) L.fastTaps = 0;
            }
            // source bounding boxes of the tiled resize's destination blocks
            L.rzBoxW = L.rzBoxH = 0;
            if (L.mode == 0 && L.fastTaps) {
                const int words = L.pitch / 4, nxb = (words + 31) / 32, Hp = L.h + 2 * ORBFE_YOFF;
                const int nyb = (Hp + ORBFE_RZ_DH - 1) / ORBFE_RZ_DH;
                L.rzXblk = tapOff;
                L.rzYblk = tapOff + (unsigned)nxb;
                taps.resize(tapOff + nxb + nyb);
                int boxW = 0, boxH = 0;
                for (int xb = 0; xb < nxb; xb++) {
                    int lo = 1 << 30, hi = 0;
                    for (int wc = 32 * xb; wc < std::min(32 * xb + 32, words); wc++)
                        for (int i = 0; i < 4; i++) {
                            int p = 4 * wc - ORBFE_XOFF + i;
                            if (p < 0) p = -p;
                            if (p >= L.w) p = 2 * (L.w - 1) - p;
                            p = std::max(0, std::min(p, L.w - 1));
                            lo = std::min(lo, (int)taps[L.xtab + p].s);
                            hi = std::max(hi, (int)taps[L.xtab + p].s1);
                        }
                    const int cLo = (ORBFE_XOFF + lo) & ~15;
                    taps[L.rzXblk + xb] = OrbfeTap{(short)cLo, 0, 0, (short)(ORBFE_XOFF + hi)};
                    boxW = std::max(boxW, ORBFE_XOFF + hi + 1 - cLo + 12);   // + the three-word window of the last thread
                }
                for (int yb = 0; yb < nyb; yb++) {
                    int lo = 1 << 30, hi = 0;
                    for (int py = ORBFE_RZ_DH * yb; py < std::min(ORBFE_RZ_DH * (yb + 1), Hp); py++) {
                        int p = py - ORBFE_YOFF;
                        if (p < 0) p = -p;
                        if (p >= L.h) p = 2 * (L.h - 1) - p;
                        p = std::max(0, std::min(p, L.h - 1));
                        lo = std::min(lo, (int)taps[L.ytab + p].s);
                        hi = std::max(hi, (int)taps[L.ytab + p].s1);
                    }
                    taps[L.rzYblk + yb] = OrbfeTap{(short)lo, 0, 0, (short)hi};
                    boxH = std::max(boxH, hi - lo + 1);
                }
                tapOff += (unsigned)(nxb + nyb);
                boxW = (boxW + 15) & ~15;
                if (boxW <= 256 && boxH <= 256) { L.rzBoxW = boxW; L.rzBoxH = boxH; }
            }
        }
        L.blurTileBase = blurTile;
        L.blurTilesX = (L.w + ORBFE_BLUR_TW - 1) / ORBFE_BLUR_TW;
        L.blurTilesY = (L.h + 4 * ORBFE_BLUR_TH - 1) / (4 * ORBFE_BLUR_TH);   // 4 warps (strips) per CTA
        blurTile += L.blurTilesX * L.blurTilesY;
    }
    g.pyrStride = off;
    g.cellsPerFrame = cell;
    g.slotsPerFrame = slot;
    g.kpCapFrame = kpBase;
    g.blurTiles = blurTile;
    if (orbfe_octree_prepare(g) < 0) return fail(ORBFE_ERR_CUDA, "cudaFuncSetAttribute(octree)", cudaGetLastError());

    // new geometry invalidates the chunk buffers
    if (int qrc = quiesce(e)) return qrc;
    if (e->d_taps) { cudaFree(e->d_taps); e->d_taps = nullptr; }
    if (!taps.empty()) {
        CK(cudaMalloc(&e->d_taps, taps.size() * sizeof(OrbfeTap)));
        CK(cudaMemcpy(e->d_taps, taps.data(), taps.size() * sizeof(OrbfeTap), cudaMemcpyHostToDevice));
    }
    if (e->d_cells) { cudaFree(e->d_cells); e->d_cells = nullptr; }
    {
        std::vector<OrbfeFastCell> cells;
        orbfe_fast_cell_table(g, cells);
        CK(cudaMalloc(&e->d_cells, cells.size() * sizeof(OrbfeFastCell)));
        CK(cudaMemcpy(e->d_cells, cells.data(), cells.size() * sizeof(OrbfeFastCell), cudaMemcpyHostToDevice));
    }
    // a captured per-frame graph holds the old tap / cell tables and chunk buffers: it dies with them
    if (e->graphExec) { cudaGraphExecDestroy(e->graphExec); e->graphExec = nullptr; }
    if (e->slab) { cudaFree(e->slab); e->slab = nullptr; }
    if (e->slab2) { cudaFree(e->slab2); e->slab2 = nullptr; }
    e->chunkCap = 0;
    e->chunkCap2 = 0;
    e->g = g;
    e->haveGeom = true;
    const size_t ocg = g.ocShared ? 0 : orbfe_octree_table_bytes(g.ocMmax) * g.nlevels;
    e->perFrameBytes = 2 * (size_t)g.pyrStride + 3 * sizeof(uint32_t) * (size_t)g.slotsPerFrame +
                       sizeof(int) * (size_t)g.cellsPerFrame + (sizeof(uint32_t) + sizeof(OrbfeWork)) * (size_t)g.kpCapFrame +
                       2 * sizeof(int) * g.nlevels + ocg + 2048;
    return ORBFE_OK;
}

int ensure_chunk_set(OrbfeExtractor* e, int frames, OrbfeChunkBufs& bufs, void*& slab, int& cap) {
    if (frames <= cap) return ORBFE_OK;
    if (int qrc = quiesce(e)) return qrc;
    if (e->graphExec) { cudaGraphExecDestroy(e->graphExec); e->graphExec = nullptr; }   // it replays the old buffers
    if (slab) { cudaFree(slab); slab = nullptr; cap = 0; }
    e->lastFrames = 0;       // nothing the taps could read any more
    e->lastBufs = nullptr;
    const OrbfeFrameGeom& g = e->g;
    const size_t B = (size_t)frames;
    size_t sz[12], total = 0;
    const size_t ocStride = g.ocShared ? 0 : orbfe_octree_table_bytes(g.ocMmax);
    sz[0] = B * g.pyrStride; sz[1] = sz[0]; sz[2] = 0;
    sz[3] = B * g.slotsPerFrame * 4; sz[4] = B * g.cellsPerFrame * 4; sz[5] = sz[3]; sz[6] = sz[3];
    sz[7] = B * g.nlevels * 4; sz[8] = B * g.kpCapFrame * 4; sz[9] = sz[7];
    sz[10] = B * g.kpCapFrame * sizeof(OrbfeWork); sz[11] = B * g.nlevels * ocStride;
    size_t offs[12];
    for (int i = 0; i < 12; i++) { offs[i] = total; total += align_up(sz[i], 256); }
    CK(cudaMalloc(&slab, total));
    char* p = (char*)slab;
    bufs.pyr = (uint8_t*)(p + offs[0]);
    bufs.blur = (uint8_t*)(p + offs[1]);
    bufs.slots = (uint32_t*)(p + offs[3]);
    bufs.cellCount = (int*)(p + offs[4]);
    bufs.cand = (uint32_t*)(p + offs[5]);
    bufs.pnode = (uint32_t*)(p + offs[6]);
    bufs.candCount = (int*)(p + offs[7]);
    bufs.kp = (uint32_t*)(p + offs[8]);
    bufs.kpCount = (int*)(p + offs[9]);
    bufs.work = (OrbfeWork*)(p + offs[10]);
    bufs.ocGlobal = ocStride ? p + offs[11] : nullptr;
    bufs.ocGlobalStride = ocStride;
    int mrc = orbfe_fast_make_maps(g, bufs, frames);
    if (mrc != ORBFE_OK) return mrc;
    if ((mrc = orbfe_resize_make_maps(g, bufs, frames)) != ORBFE_OK) return mrc;
    cap = frames;
    return ORBFE_OK;
}

int ensure_chunk(OrbfeExtractor* e, int frames) { return ensure_chunk_set(e, frames, e->bufs, e->slab, e->chunkCap); }

// The second set, used by the odd chunks of a multi-chunk host batch.
int ensure_chunk2(OrbfeExtractor* e, int frames) {
    if (!e->sCompute2) CK(cudaStreamCreateWithFlags(&e->sCompute2, cudaStreamNonBlocking));
    return ensure_chunk_set(e, frames, e->bufs2, e->slab2, e->chunkCap2);
}

constexpr int kGraphMaxFrames = 4;   // host-pointer calls of up to this many frames replay a captured CUDA graph

int chunk_frames(const OrbfeExtractor* e, int B) {
    size_t c = e->maxBytes / std::max<size_t>(e->perFrameBytes, 1);
    c = std::max<size_t>(1, std::min<size_t>(c, 4096));
    return (int)std::min<size_t>(c, (size_t)B);
}

void stage_mark(OrbfeExtractor* e, int i, cudaStream_t st) {
    if (e->profiling) cudaEventRecord(e->evStage[e->profCount % OrbfeExtractor::kProfSets][i], st);
}

// Enqueue the whole extraction of `B` (<= chunkCap) frames on `st`.
void enqueue_chunk(OrbfeExtractor* e, const uint8_t* d_images, size_t step, size_t frameStride, int B,
                   int lap0, int lap1, OrbfeKeyPoint* d_kps, uint8_t* d_desc, int capacity, int* d_n,
                   int* d_mono, cudaStream_t st, const OrbfeChunkBufs* set = nullptr) {
    const OrbfeFrameGeom& g = e->g;
    const OrbfeChunkBufs& bufs = set ? *set : e->bufs;
    OrbfeRectify rect;
    rect.mapx = e->d_mapx; rect.mapy = e->d_mapy; rect.srcRows = e->srcRows; rect.srcCols = e->srcCols;
    stage_mark(e, 1, st);
    orbfe_launch_pyramid(g, e->d_taps, d_images, step, frameStride, bufs, B, st, &e->launches, e->d_mapx ? &rect : nullptr);
    stage_mark(e, 2, st);
    orbfe_launch_fast(g, e->d_cells, bufs, B, st, &e->launches);
    stage_mark(e, 3, st);
    orbfe_launch_octree(g, bufs, B, st, &e->launches);
    stage_mark(e, 4, st);
    orbfe_launch_layout(g, bufs, B, lap0, lap1, d_kps, capacity, d_n, d_mono, st, &e->launches);
    stage_mark(e, 5, st);
    orbfe_launch_blur(g, bufs, B, st, &e->launches);
    stage_mark(e, 6, st);
    orbfe_launch_describe(g, bufs, B, d_kps, d_desc, capacity, st, &e->launches);
    stage_mark(e, 7, st);
    if (e->profiling) e->profCount++;
    e->lastFrames = B;
    e->lastBufs = &bufs;     // the taps (mvImagePyramid, stage taps, stereo matcher) read the set the last chunk used
}

int ensure_staging(OrbfeExtractor* e, int frames, int rows, int cols, int capacity) {
    const size_t need = (size_t)frames * rows * cols;
    if (need > e->inBytes) {
        if (int qrc = quiesce(e)) return qrc;
        if (e->graphExec) { cudaGraphExecDestroy(e->graphExec); e->graphExec = nullptr; }
        for (int s = 0; s < 2; s++) {
            if (e->d_in[s]) cudaFree(e->d_in[s]);
            e->d_in[s] = nullptr;
            CK(cudaMalloc(&e->d_in[s], need));
        }
        e->inBytes = need;
    }
    const size_t elems = (size_t)frames * capacity;
    if (elems > e->outElems || frames > e->outFrames) {
        if (int qrc = quiesce(e)) return qrc;
        if (e->graphExec) { cudaGraphExecDestroy(e->graphExec); e->graphExec = nullptr; }
        for (int s = 0; s < 2; s++) {
            if (e->d_okps[s]) cudaFree(e->d_okps[s]);
            if (e->d_odesc[s]) cudaFree(e->d_odesc[s]);
            if (e->d_on[s]) cudaFree(e->d_on[s]);
            if (e->d_omono[s]) cudaFree(e->d_omono[s]);
            e->d_okps[s] = nullptr; e->d_odesc[s] = nullptr; e->d_on[s] = nullptr; e->d_omono[s] = nullptr;
            CK(cudaMalloc(&e->d_okps[s], elems * sizeof(OrbfeKeyPoint)));
            CK(cudaMalloc(&e->d_odesc[s], elems * 32));
            CK(cudaMalloc(&e->d_on[s], (size_t)frames * sizeof(int)));
            CK(cudaMalloc(&e->d_omono[s], (size_t)frames * sizeof(int)));
        }
        e->outFrames = frames;
        e->outElems = elems;
    }
    return ORBFE_OK;
}

int check_handle(const OrbfeExtractor* h) {
    if (!h) return fail(ORBFE_ERR_INVALID, "null extractor");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(ORBFE_ERR_CUDA, "cudaSetDevice", e);
    return ORBFE_OK;
}

}  // namespace

extern "C" {

const char* orbfe_last_error(void) { return g_err.c_str(); }

const char* orbfe_version(void) { return "orbfe-b200 sm_100a " __DATE__; }

int orbfe_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                           int device, OrbfeExtractor** out) {
    if (!out) return fail(ORBFE_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (nfeatures < 0 || nlevels < 1 || nlevels > ORBFE_MAX_LEVELS || !(scaleFactor > 0.f))
        return fail(ORBFE_ERR_INVALID, "bad extractor parameters");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return fail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return fail(ORBFE_ERR_INVALID, "bad device ordinal");
    CK(cudaSetDevice(device));
    OrbfeExtractor* e = new OrbfeExtractor();
    e->nfeatures = nfeatures;
    e->scaleFactor = scaleFactor;
    e->nlevels = nlevels;
    e->iniTh = iniThFAST;
    e->minTh = minThFAST;
    e->device = device;
    build_tables(e);
    if (const char* mb = getenv("ORBFE_MAX_BYTES")) e->maxBytes = std::max<size_t>((size_t)strtoull(mb, nullptr, 10), (size_t)64 << 20);
    cudaError_t er = cudaStreamCreateWithFlags(&e->sCompute, cudaStreamNonBlocking);
    if (er == cudaSuccess) er = cudaStreamCreateWithFlags(&e->sH2D, cudaStreamNonBlocking);
    if (er == cudaSuccess) er = cudaStreamCreateWithFlags(&e->sD2H, cudaStreamNonBlocking);
    for (int s = 0; s < 2 && er == cudaSuccess; s++) {
        er = cudaEventCreateWithFlags(&e->evIn[s], cudaEventDisableTiming);
        if (er == cudaSuccess) er = cudaEventCreateWithFlags(&e->evInFree[s], cudaEventDisableTiming);
        if (er == cudaSuccess) er = cudaEventCreateWithFlags(&e->evDone[s], cudaEventDisableTiming);
        if (er == cudaSuccess) er = cudaEventCreateWithFlags(&e->evOutFree[s], cudaEventDisableTiming);
    }
    for (int k = 0; k < OrbfeExtractor::kProfSets && er == cudaSuccess; k++)
        for (int i = 0; i <= ORBFE_NUM_STAGES && er == cudaSuccess; i++) er = cudaEventCreate(&e->evStage[k][i]);
    if (er != cudaSuccess) {
        orbfe_extractor_destroy(e);
        return fail(ORBFE_ERR_CUDA, "stream/event creation", er);
    }
    *out = e;
    return ORBFE_OK;
}

void orbfe_extractor_destroy(OrbfeExtractor* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->sCompute) cudaStreamSynchronize(e->sCompute);
    if (e->sCompute2) cudaStreamSynchronize(e->sCompute2);
    if (e->sD2H) cudaStreamSynchronize(e->sD2H);
    if (e->sH2D) cudaStreamSynchronize(e->sH2D);
    if (e->graphExec) { cudaGraphExecDestroy(e->graphExec); e->graphExec = nullptr; }
    if (e->d_mapx) cudaFree(e->d_mapx);
    if (e->d_mapy) cudaFree(e->d_mapy);
    if (e->slab) cudaFree(e->slab);
    if (e->slab2) cudaFree(e->slab2);
    if (e->d_taps) cudaFree(e->d_taps);
    if (e->d_cells) cudaFree(e->d_cells);
    if (e->d_stereoSad) cudaFree(e->d_stereoSad);
    for (int s = 0; s < 2; s++) {
        if (e->d_in[s]) cudaFree(e->d_in[s]);
        if (e->d_okps[s]) cudaFree(e->d_okps[s]);
        if (e->d_odesc[s]) cudaFree(e->d_odesc[s]);
        if (e->d_on[s]) cudaFree(e->d_on[s]);
        if (e->d_omono[s]) cudaFree(e->d_omono[s]);
        if (e->evIn[s]) cudaEventDestroy(e->evIn[s]);
        if (e->evInFree[s]) cudaEventDestroy(e->evInFree[s]);
        if (e->evDone[s]) cudaEventDestroy(e->evDone[s]);
        if (e->evOutFree[s]) cudaEventDestroy(e->evOutFree[s]);
    }
    for (int s = 0; s < OrbfeExtractor::kMaxPending; s++) {
        if (e->evPending[s]) cudaEventDestroy(e->evPending[s]);
    }
    for (int k = 0; k < OrbfeExtractor::kProfSets; k++)
        for (int i = 0; i <= ORBFE_NUM_STAGES; i++)
            if (e->evStage[k][i]) cudaEventDestroy(e->evStage[k][i]);
    if (e->sCompute) cudaStreamDestroy(e->sCompute);
    if (e->sCompute2) cudaStreamDestroy(e->sCompute2);
    if (e->sH2D) cudaStreamDestroy(e->sH2D);
    if (e->sD2H) cudaStreamDestroy(e->sD2H);
    delete e;
}

int orbfe_get_levels(const OrbfeExtractor* h) { return h ? h->nlevels : ORBFE_ERR_INVALID; }
float orbfe_get_scale_factor(const OrbfeExtractor* h) { return h ? (float)h->scaleFactor : 0.f; }

int orbfe_scale_tables(const OrbfeExtractor* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
    if (!h) return fail(ORBFE_ERR_INVALID, "null extractor");
    for (int i = 0; i < h->nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->invScale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->invSigma2[i];
    }
    return ORBFE_OK;
}

int orbfe_features_per_level(const OrbfeExtractor* h, int* n_per_level) {
    if (!h || !n_per_level) return fail(ORBFE_ERR_INVALID, "null argument");
    for (int i = 0; i < h->nlevels; i++) n_per_level[i] = h->nfeat[i];
    return ORBFE_OK;
}

int orbfe_max_keypoints(const OrbfeExtractor* h) {
    if (!h) return ORBFE_ERR_INVALID;
    int n = 0;  // per level: DistributeOctTree stops at >= N after adding <= 3 nodes, or at 4 per root (<= 8 roots here)
    for (int i = 0; i < h->nlevels; i++) n += std::max(h->nfeat[i] + 3, 4 * 8) + 1;
    return n;
}

// The same bound with the root count of the actual frame size: a level of a very wide frame starts from
// round(width / height) roots, each of which can leave four nodes behind even when the level's target is smaller.
int orbfe_max_keypoints_for(const OrbfeExtractor* h, int rows, int cols) {
    if (!h || rows <= 0 || cols <= 0) return ORBFE_ERR_INVALID;
    if (h->d_mapx) { rows = h->rectRows; cols = h->rectCols; }   // the frames are rectified to this size first
    int n = 0;
    for (int l = 0; l < h->nlevels; l++) n += std::max(h->nfeat[l] + 3, 4 * std::max(level_roots(h, rows, cols, l), 1)) + 1;
    return n;
}

int orbfe_level_size(const OrbfeExtractor* h, int rows, int cols, int level, int* w, int* hgt) {
    if (!h || level < 0 || level >= h->nlevels) return fail(ORBFE_ERR_INVALID, "bad level");
    if (w) *w = cv_round((float)cols * h->invScale[level]);
    if (hgt) *hgt = cv_round((float)rows * h->invScale[level]);
    return ORBFE_OK;
}

int orbfe_extractor_set_rectification(OrbfeExtractor* h, const float* map_x, const float* map_y, int rows, int cols) {
    int rc = check_handle(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->sCompute));
    if (h->sCompute2) CK(cudaStreamSynchronize(h->sCompute2));
    if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }
    if (h->d_mapx) { cudaFree(h->d_mapx); h->d_mapx = nullptr; }
    if (h->d_mapy) { cudaFree(h->d_mapy); h->d_mapy = nullptr; }
    h->rectRows = h->rectCols = 0;
    if (!map_x && !map_y) return ORBFE_OK;   // rectification off
    if (!map_x || !map_y || rows <= 0 || cols <= 0) return fail(ORBFE_ERR_INVALID, "rectification maps: bad arguments");
    const size_t bytes = sizeof(float) * (size_t)rows * cols;
    CK(cudaMalloc(&h->d_mapx, bytes));
    CK(cudaMalloc(&h->d_mapy, bytes));
    CK(cudaMemcpy(h->d_mapx, map_x, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_mapy, map_y, bytes, cudaMemcpyHostToDevice));
    h->rectRows = rows; h->rectCols = cols;
    return ORBFE_OK;
}

int orbfe_set_max_bytes(OrbfeExtractor* h, unsigned long long bytes) {
    if (!h) return fail(ORBFE_ERR_INVALID, "null extractor");
    h->maxBytes = std::max<size_t>((size_t)bytes, (size_t)64 << 20);
    return ORBFE_OK;
}

int orbfe_extract_batch_device(OrbfeExtractor* h, const uint8_t* d_images, int B, int rows, int cols,
                               size_t step, size_t frame_stride, int lap0, int lap1,
                               OrbfeKeyPoint* d_keypoints, uint8_t* d_descriptors, int capacity,
                               int* d_n_out, int* d_mono_out, void* stream) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!d_images || rows <= 0 || cols <= 0) return fail(ORBFE_EMPTY_IMAGE, "empty image");
    if (B <= 0 || capacity <= 0 || step < (size_t)cols || !d_keypoints || !d_descriptors || !d_n_out || !d_mono_out)
        return fail(ORBFE_ERR_INVALID, "bad batch arguments");
    if (h->pendCount > 0) return fail(ORBFE_ERR_INVALID, "host batches in flight (orbfe_extract_batch_wait first)");
    h->srcRows = rows; h->srcCols = cols;
    if ((rc = build_geometry(h, h->d_mapx ? h->rectRows : rows, h->d_mapx ? h->rectCols : cols))) return rc;
    const int chunk = chunk_frames(h, B);
    if ((rc = ensure_chunk(h, chunk))) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->sCompute;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        enqueue_chunk(h, d_images + (size_t)b0 * frame_stride, step, frame_stride, nb, lap0, lap1,
                      d_keypoints + (size_t)b0 * capacity, d_descriptors + (size_t)b0 * capacity * 32,
                      capacity, d_n_out + b0, d_mono_out + b0, st);
    }
    CK(cudaGetLastError());
    return ORBFE_OK;
}

// Enqueue one host batch: chunked H2D / kernels / D2H over the double-buffered staging slots, completion event on
// sD2H.  Does not wait.  `h->pendCount > 0` on entry means the pipeline is already full (an earlier submit is still
// running): the slots keep alternating from where that batch left them and the fill ramp is skipped.
static int submit_host_batch(OrbfeExtractor* h, const uint8_t* images, int B, int rows, int cols, size_t step,
                             size_t frame_stride, int lap0, int lap1, OrbfeKeyPoint* keypoints,
                             uint8_t* descriptors, int capacity, int* n_out, int* mono_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!images || rows <= 0 || cols <= 0) return fail(ORBFE_EMPTY_IMAGE, "empty image");
    if (B <= 0 || capacity <= 0 || step < (size_t)cols || !keypoints || !descriptors || !n_out || !mono_out)
        return fail(ORBFE_ERR_INVALID, "bad batch arguments");
    if (h->pendCount >= OrbfeExtractor::kMaxPending) return fail(ORBFE_ERR_INVALID, "too many host batches in flight");
    h->srcRows = rows; h->srcCols = cols;
    if ((rc = build_geometry(h, h->d_mapx ? h->rectRows : rows, h->d_mapx ? h->rectCols : cols))) return rc;
    const int chunk = chunk_frames(h, B);
    if ((rc = ensure_chunk(h, chunk))) return rc;
    // more than one chunk: odd chunks run on a second stream with their own intermediates (not while profiling:
    // the stage events belong to one stream)
    const bool streaming = h->pendCount > 0;
    const bool dual = (B > chunk || streaming) && !h->profiling;
    if (dual && (rc = ensure_chunk2(h, chunk))) return rc;
    if ((rc = ensure_staging(h, chunk, rows, cols, capacity))) return rc;
    if (!streaming) h->chunkSeq = 0;   // idle pipeline: every event of earlier calls has completed
    const size_t fbytes = (size_t)rows * cols;
    const bool packed = step == (size_t)cols && frame_stride == fbytes;
    int ci = 0, nb = 0;
    // a failure from here on leaves chunks in flight that nobody will wait for: drain the handle before reporting it
#define CKQ(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) {                                           \
            cudaStreamCaptureStatus cs_ = cudaStreamCaptureStatusNone;     \
            for (cudaStream_t s_ : {h->sCompute, h->sCompute2}) {          \
                cudaGraph_t g_ = nullptr;                                  \
                if (s_ && cudaStreamIsCapturing(s_, &cs_) == cudaSuccess && cs_ != cudaStreamCaptureStatusNone) { \
                    cudaStreamEndCapture(s_, &g_);                         \
                    if (g_) cudaGraphDestroy(g_);                          \
                }                                                          \
            }                                                              \
            quiesce(h);                                                    \
            return fail(ORBFE_ERR_CUDA, #call, e_);                        \
        }                                                                  \
    } while (0)
    for (int b0 = 0; b0 < B; b0 += nb, ci++, h->chunkSeq++) {
        // Pipeline fill: the H2D of the first chunk overlaps nothing, and the kernels of chunk k cannot start before
        // the H2D of chunk k has landed, so chunk k+1 must not take longer to copy (~7 us per frame over PCIe) than
        // chunk k takes to compute (~12 us per frame): chunk/4, chunk/2, then full chunks.
        nb = (ci < 2 && B > chunk && !streaming) ? std::max(chunk >> (2 - ci), 1) : chunk;
        nb = std::min(nb, B - b0);
        const int s = (int)(h->chunkSeq & 1);
        const bool reuse = h->chunkSeq >= 2;   // the slot has been used since the pipeline was last idle
        // H2D of this chunk overlaps the kernels of the previous one
        if (reuse) CKQ(cudaStreamWaitEvent(h->sH2D, h->evInFree[s], 0));
        const uint8_t* src = images + (size_t)b0 * frame_stride;
        if (packed) {
            CKQ(cudaMemcpyAsync(h->d_in[s], src, fbytes * nb, cudaMemcpyHostToDevice, h->sH2D));
        } else {
            for (int f = 0; f < nb; f++)
                CKQ(cudaMemcpy2DAsync(h->d_in[s] + f * fbytes, cols, src + (size_t)f * frame_stride, step,
                                     cols, rows, cudaMemcpyHostToDevice, h->sH2D));
        }
        CKQ(cudaEventRecord(h->evIn[s], h->sH2D));
        cudaStream_t sc = (dual && s) ? h->sCompute2 : h->sCompute;
        CKQ(cudaStreamWaitEvent(sc, h->evIn[s], 0));
        if (reuse) CKQ(cudaStreamWaitEvent(sc, h->evOutFree[s], 0));
        if (B <= kGraphMaxFrames && B <= chunk && !h->profiling && !streaming) {
            // per-frame call: one graph launch instead of 15 kernel launches
            OrbfeExtractor::GraphKey key;
            key.rows = rows; key.cols = cols; key.B = nb; key.lap0 = lap0; key.lap1 = lap1; key.capacity = capacity;
            key.in = h->d_in[s]; key.kps = h->d_okps[s]; key.desc = h->d_odesc[s]; key.slab = h->slab; key.map = h->d_mapx;
            if (!h->graphExec || !(key == h->graphKey)) {
                if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }
                cudaGraph_t graph = nullptr;
                const long long l0 = h->launches;
                CKQ(cudaStreamBeginCapture(sc, cudaStreamCaptureModeThreadLocal));
                enqueue_chunk(h, h->d_in[s], cols, fbytes, nb, lap0, lap1, h->d_okps[s], h->d_odesc[s], capacity,
                              h->d_on[s], h->d_omono[s], sc, nullptr);
                CKQ(cudaStreamEndCapture(sc, &graph));
                h->graphLaunches = h->launches - l0;
                h->launches = l0;
                cudaError_t ge = cudaGraphInstantiate(&h->graphExec, graph, 0);
                cudaGraphDestroy(graph);
                if (ge != cudaSuccess) { h->graphExec = nullptr; return fail(ORBFE_ERR_CUDA, "cudaGraphInstantiate", ge); }
                h->graphKey = key;
            }
            CKQ(cudaGraphLaunch(h->graphExec, sc));
            h->launches += h->graphLaunches;
            h->lastFrames = nb;
            h->lastBufs = &h->bufs;
        } else {
            enqueue_chunk(h, h->d_in[s], cols, fbytes, nb, lap0, lap1, h->d_okps[s], h->d_odesc[s], capacity,
                          h->d_on[s], h->d_omono[s], sc, (dual && s) ? &h->bufs2 : nullptr);
        }
        CKQ(cudaEventRecord(h->evInFree[s], sc));
        CKQ(cudaEventRecord(h->evDone[s], sc));
        CKQ(cudaStreamWaitEvent(h->sD2H, h->evDone[s], 0));
        CKQ(cudaMemcpyAsync(keypoints + (size_t)b0 * capacity, h->d_okps[s], (size_t)nb * capacity * sizeof(OrbfeKeyPoint),
                           cudaMemcpyDeviceToHost, h->sD2H));
        CKQ(cudaMemcpyAsync(descriptors + (size_t)b0 * capacity * 32, h->d_odesc[s], (size_t)nb * capacity * 32,
                           cudaMemcpyDeviceToHost, h->sD2H));
        CKQ(cudaMemcpyAsync(n_out + b0, h->d_on[s], nb * sizeof(int), cudaMemcpyDeviceToHost, h->sD2H));
        CKQ(cudaMemcpyAsync(mono_out + b0, h->d_omono[s], nb * sizeof(int), cudaMemcpyDeviceToHost, h->sD2H));
        CKQ(cudaEventRecord(h->evOutFree[s], h->sD2H));
    }
#undef CKQ
    const int slot = (h->pendHead + h->pendCount) % OrbfeExtractor::kMaxPending;
    if (!h->evPending[slot]) CK(cudaEventCreateWithFlags(&h->evPending[slot], cudaEventDisableTiming));
    CK(cudaEventRecord(h->evPending[slot], h->sD2H));
    h->pending[slot].n_out = n_out; h->pending[slot].B = B; h->pending[slot].capacity = capacity;
    h->pendCount++;
    CK(cudaGetLastError());
    return ORBFE_OK;
}

// Wait for the oldest batch in flight (its last D2H is behind everything else it enqueued).
static int wait_host_batch(OrbfeExtractor* h) {
    const int slot = h->pendHead;
    const OrbfeExtractor::Pending p = h->pending[slot];
    h->pendHead = (h->pendHead + 1) % OrbfeExtractor::kMaxPending;
    h->pendCount--;
    CK(cudaEventSynchronize(h->evPending[slot]));
    CK(cudaGetLastError());
    for (int b = 0; b < p.B; b++)
        if (p.n_out[b] > p.capacity) return fail(ORBFE_ERR_CAPACITY, "capacity smaller than the keypoint count (see orbfe_max_keypoints)");
    return ORBFE_OK;
}

int orbfe_extract_batch_submit(OrbfeExtractor* h, const uint8_t* images, int B, int rows, int cols, size_t step,
                               size_t frame_stride, int lap0, int lap1, OrbfeKeyPoint* keypoints,
                               uint8_t* descriptors, int capacity, int* n_out, int* mono_out) {
    return submit_host_batch(h, images, B, rows, cols, step, frame_stride, lap0, lap1, keypoints, descriptors, capacity,
                             n_out, mono_out);
}

int orbfe_extract_batch_wait(OrbfeExtractor* h) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (h->pendCount <= 0) return fail(ORBFE_ERR_INVALID, "no host batch in flight");
    return wait_host_batch(h);
}

int orbfe_extract_batch(OrbfeExtractor* h, const uint8_t* images, int B, int rows, int cols, size_t step,
                        size_t frame_stride, int lap0, int lap1, OrbfeKeyPoint* keypoints,
                        uint8_t* descriptors, int capacity, int* n_out, int* mono_out) {
    int rc = submit_host_batch(h, images, B, rows, cols, step, frame_stride, lap0, lap1, keypoints, descriptors,
                               capacity, n_out, mono_out);
    if (rc) return rc;
    // synchronous call: drain everything in flight, this batch last; the first error wins
    int first = ORBFE_OK;
    while (h->pendCount > 0) {
        rc = wait_host_batch(h);
        if (first == ORBFE_OK) first = rc;
    }
    CK(cudaStreamSynchronize(h->sCompute));
    if (h->sCompute2) CK(cudaStreamSynchronize(h->sCompute2));
    return first;
}

int orbfe_extract(OrbfeExtractor* h, const uint8_t* image, int rows, int cols, size_t step, int lap0,
                  int lap1, OrbfeKeyPoint* keypoints, uint8_t* descriptors, int capacity, int* n_out) {
    int n = 0, mono = 0;
    if (n_out) *n_out = 0;
    const int rc = orbfe_extract_batch(h, image, 1, rows, cols, step, (size_t)rows * step, lap0, lap1,
                                       keypoints, descriptors, capacity, &n, &mono);
    if (n_out) *n_out = n;
    if (rc != ORBFE_OK) return rc;
    return mono;  // ORBextractor.cc:1681
}

// ---- mvImagePyramid and the stage taps ----------------------------------------------------
static int tap_check(OrbfeExtractor* h, int frame, int level) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!h->haveGeom || frame < 0 || frame >= h->lastFrames || level < 0 || level >= h->nlevels)
        return fail(ORBFE_ERR_INVALID, "no such frame/level in the last extract call");
    CK(cudaStreamSynchronize(h->sCompute));
    if (h->sCompute2) CK(cudaStreamSynchronize(h->sCompute2));   // odd chunks of a multi-chunk batch ran there
    return ORBFE_OK;
}

static int copy_level(OrbfeExtractor* h, const uint8_t* slab, int frame, int level, int with_border,
                      uint8_t* dst, size_t dst_step) {
    const OrbfeLevelGeom& L = h->g.lv[level];
    const uint8_t* base = slab + (size_t)frame * h->g.pyrStride + L.off;
    if (with_border) {
        CK(cudaMemcpy2D(dst, dst_step, base + ORBFE_XOFF - ORBFE_EDGE, L.pitch, L.w + 2 * ORBFE_EDGE,
                        L.h + 2 * ORBFE_EDGE, cudaMemcpyDeviceToHost));
    } else {
        CK(cudaMemcpy2D(dst, dst_step, base + (size_t)ORBFE_YOFF * L.pitch + ORBFE_XOFF, L.pitch, L.w, L.h,
                        cudaMemcpyDeviceToHost));
    }
    return ORBFE_OK;
}

int orbfe_pyramid_level(OrbfeExtractor* h, int frame, int level, int with_border, uint8_t* dst, size_t dst_step) {
    int rc = tap_check(h, frame, level);
    if (rc) return rc;
    return copy_level(h, (h->lastBufs ? h->lastBufs : &h->bufs)->pyr, frame, level, with_border, dst, dst_step);
}

int orbfe_debug_blurred(OrbfeExtractor* h, int frame, int level, uint8_t* dst, size_t dst_step) {
    int rc = tap_check(h, frame, level);
    if (rc) return rc;
    return copy_level(h, (h->lastBufs ? h->lastBufs : &h->bufs)->blur, frame, level, 0, dst, dst_step);
}

int orbfe_debug_candidates(OrbfeExtractor* h, int frame, int level, int32_t* xys, int capacity, int* n_out) {
    int rc = tap_check(h, frame, level);
    if (rc) return rc;
    const OrbfeFrameGeom& g = h->g;
    const OrbfeLevelGeom& L = g.lv[level];
    int n = 0;
    const OrbfeChunkBufs& lb = h->lastBufs ? *h->lastBufs : h->bufs;
    CK(cudaMemcpy(&n, lb.candCount + frame * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    if (n_out) *n_out = n;
    const int m = std::min(n, capacity);
    std::vector<uint32_t> pk(std::max(m, 1));
    if (m) CK(cudaMemcpy(pk.data(), lb.cand + (size_t)frame * g.slotsPerFrame + L.slotBase, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) {
        xys[3 * i] = (int)(pk[i] & 0xFFF);
        xys[3 * i + 1] = (int)((pk[i] >> 12) & 0xFFF);
        xys[3 * i + 2] = (int)(pk[i] >> 24);
    }
    return ORBFE_OK;
}

int orbfe_debug_level_keypoints(OrbfeExtractor* h, int frame, int level, int32_t* xys, int capacity, int* n_out) {
    int rc = tap_check(h, frame, level);
    if (rc) return rc;
    const OrbfeFrameGeom& g = h->g;
    const OrbfeLevelGeom& L = g.lv[level];
    int n = 0;
    const OrbfeChunkBufs& lb = h->lastBufs ? *h->lastBufs : h->bufs;
    CK(cudaMemcpy(&n, lb.kpCount + frame * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    if (n_out) *n_out = n;
    const int m = std::min(n, capacity);
    std::vector<uint32_t> pk(std::max(m, 1));
    if (m) CK(cudaMemcpy(pk.data(), lb.kp + (size_t)frame * g.kpCapFrame + L.kpBase, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) {
        xys[3 * i] = (int)(pk[i] & 0xFFF);
        xys[3 * i + 1] = (int)((pk[i] >> 12) & 0xFFF);
        xys[3 * i + 2] = (int)(pk[i] >> 24);
    }
    return ORBFE_OK;
}

int orbfe_debug_octree(OrbfeExtractor* h, const int32_t* xys, int n, int minX, int maxX, int minY, int maxY,
                       int N, int32_t* keep_idx, int capacity, int* n_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_out) *n_out = 0;
    if (n <= 0) return ORBFE_OK;
    const int width = maxX - minX, height = maxY - minY;
    if (width <= 0 || height <= 0 || width > 4095 || height > 4095 || N < 0) return fail(ORBFE_ERR_INVALID, "bad octree window");
    const int nIni = (int)roundf((float)width / (float)height);
    if (nIni < 1) return fail(ORBFE_ERR_INVALID, "aspect ratio < 0.5");
    const float hX = (float)width / nIni;
    const int M = std::max(N + 3, 4 * nIni) + 1;
    std::vector<uint32_t> pk(n);
    for (int i = 0; i < n; i++)
        pk[i] = ((uint32_t)xys[3 * i + 2] << 24) | ((uint32_t)xys[3 * i + 1] << 12) | (uint32_t)xys[3 * i];
    uint32_t *d_pk = nullptr, *d_pn = nullptr;
    int* d_out = nullptr;
    char* d_tab = nullptr;
    CK(cudaMalloc(&d_pk, sizeof(uint32_t) * n));
    CK(cudaMalloc(&d_pn, sizeof(uint32_t) * n));
    CK(cudaMalloc(&d_out, sizeof(int) * (M + 1)));
    CK(cudaMalloc(&d_tab, orbfe_octree_table_bytes(M)));
    CK(cudaMemcpy(d_pk, pk.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    orbfe_launch_octree_debug(d_pk, d_pn, n, width, height, nIni, hX, N, M, d_out + 1, d_out, d_tab, h->sCompute);
    cudaError_t ce = cudaStreamSynchronize(h->sCompute);
    std::vector<int> out(M + 1, 0);
    if (ce == cudaSuccess) ce = cudaMemcpy(out.data(), d_out, sizeof(int) * (M + 1), cudaMemcpyDeviceToHost);
    cudaFree(d_pk); cudaFree(d_pn); cudaFree(d_out); cudaFree(d_tab);
    if (ce != cudaSuccess) return fail(ORBFE_ERR_CUDA, "octree debug kernel", ce);
    if (n_out) *n_out = out[0];
    for (int i = 0; i < out[0] && i < capacity; i++) keep_idx[i] = out[1 + i];
    return ORBFE_OK;
}

int orbfe_set_profiling(OrbfeExtractor* h, int enable) {
    if (!h) return fail(ORBFE_ERR_INVALID, "null extractor");
    h->profiling = enable != 0;
    h->profCount = 0;
    return ORBFE_OK;
}

// Mean per-kernel milliseconds over the chunks enqueued since orbfe_set_profiling(h, 1) (at most
// the last 64), CUDA events recorded on the launching stream around every kernel.
int orbfe_stage_ms(OrbfeExtractor* h, float* ms) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!ms) return fail(ORBFE_ERR_INVALID, "null ms");
    for (int i = 0; i < ORBFE_NUM_STAGES; i++) ms[i] = 0.f;
    const int sets = std::min(h->profCount, (int)OrbfeExtractor::kProfSets);
    if (!h->profiling || sets == 0) return fail(ORBFE_ERR_INVALID, "no profiled call since orbfe_set_profiling(h, 1)");
    for (int k = 0; k < sets; k++) {
        CK(cudaEventSynchronize(h->evStage[k][ORBFE_NUM_STAGES - 1]));
        for (int i = 1; i <= ORBFE_NUM_STAGES - 2; i++) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, h->evStage[k][i], h->evStage[k][i + 1]));
            ms[i] += t / sets;
        }
    }
    return sets;
}

long long orbfe_launch_count(const OrbfeExtractor* h) { return h ? h->launches : 0; }

// Frame::ComputeStereoMatches (src/Frame.cc:1102-1358) on the pyramids both extractors hold.
int orbfe_stereo_match(OrbfeExtractor* left, OrbfeExtractor* right, int frame, const OrbfeKeyPoint* keys_l,
                       const uint8_t* desc_l, int nl, const OrbfeKeyPoint* keys_r, const uint8_t* desc_r, int nr,
                       float mbf, float mb, float* u_right, float* depth) {
    int rc = check_handle(left);
    if (rc) return rc;
    if (!right || right->device != left->device) return fail(ORBFE_ERR_INVALID, "left/right extractors must live on one device");
    if (!left->haveGeom || !right->haveGeom || left->g.rows != right->g.rows || left->g.cols != right->g.cols ||
        left->nlevels != right->nlevels || left->scaleFactor != right->scaleFactor)
        return fail(ORBFE_ERR_INVALID, "left/right extractors hold different pyramid geometries");
    if (frame < 0 || frame >= left->lastFrames || frame >= right->lastFrames)
        return fail(ORBFE_ERR_INVALID, "no such frame in the last extract calls");
    if (nl < 0 || nr < 0 || nr >= 65536) return fail(ORBFE_ERR_INVALID, "bad keypoint counts (nr must be < 65536)");
    if (nl == 0) return ORBFE_OK;
    if (!keys_l || !desc_l || !u_right || !depth || (nr > 0 && (!keys_r || !desc_r))) return fail(ORBFE_ERR_INVALID, "null argument");
    CK(cudaStreamSynchronize(right->sCompute));
    CK(cudaStreamSynchronize(left->sCompute));
    if (right->sCompute2) CK(cudaStreamSynchronize(right->sCompute2));
    if (left->sCompute2) CK(cudaStreamSynchronize(left->sCompute2));
    const OrbfeChunkBufs& lbL = left->lastBufs ? *left->lastBufs : left->bufs;
    const OrbfeChunkBufs& lbR = right->lastBufs ? *right->lastBufs : right->bufs;
    OrbfeStage S;
    const size_t ikl = S.in(keys_l, sizeof(OrbfeKeyPoint) * (size_t)nl), idl = S.in(desc_l, 32 * (size_t)nl);
    const size_t ikr = S.in(keys_r, sizeof(OrbfeKeyPoint) * (size_t)nr), idr = S.in(desc_r, 32 * (size_t)nr);
    const size_t wsad = S.work(4 * (size_t)nl);
    const size_t widx = S.work(4 * orbfe_stereo_index_ints(left->g, 1, nr));   // candidate index per band of rows
    const size_t our = S.out(u_right, 4 * (size_t)nl), odp = S.out(depth, 4 * (size_t)nl);
    CK(S.commit(left->device));
    cudaError_t e = S.upload();
    if (e == cudaSuccess) {
        orbfe_launch_stereo(left->g, lbL.pyr + (size_t)frame * left->g.pyrStride,
                            lbR.pyr + (size_t)frame * right->g.pyrStride, S.ptr<OrbfeKeyPoint>(ikl),
                            S.ptr<uint32_t>(idl), nl, S.ptr<OrbfeKeyPoint>(ikr), S.ptr<uint32_t>(idr), nr, mbf, mb,
                            S.ptr<float>(our), S.ptr<float>(odp), S.ptr<int>(wsad), S.ptr<int>(widx), S.stream());
        left->launches += 3;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = S.download();
    if (e != cudaSuccess) return fail(ORBFE_ERR_CUDA, "stereo match", e);
    return ORBFE_OK;
}

// Frame::ComputeStereoMatches for a BATCH of rectified pairs that never leaves the device: the pyramids both extractors
// hold from their last orbfe_extract_batch_device call (one chunk: frames 0 .. B-1) and that call's output slabs.
int orbfe_stereo_match_batch_device(OrbfeExtractor* left, OrbfeExtractor* right, int B, const OrbfeKeyPoint* d_keys_l,
                                    const uint8_t* d_desc_l, const int* d_n_l, const OrbfeKeyPoint* d_keys_r,
                                    const uint8_t* d_desc_r, const int* d_n_r, int capacity, float mbf, float mb,
                                    float* d_u_right, float* d_depth, void* stream) {
    int rc = check_handle(left);
    if (rc) return rc;
    if (!right || right->device != left->device) return fail(ORBFE_ERR_INVALID, "left/right extractors must live on one device");
    if (!left->haveGeom || !right->haveGeom || left->g.rows != right->g.rows || left->g.cols != right->g.cols ||
        left->nlevels != right->nlevels || left->scaleFactor != right->scaleFactor)
        return fail(ORBFE_ERR_INVALID, "left/right extractors hold different pyramid geometries");
    if (B <= 0 || B > left->lastFrames || B > right->lastFrames || B > left->chunkCap || B > right->chunkCap)
        return fail(ORBFE_ERR_INVALID, "the pairs must be the frames of the last (single-chunk) extract call of both extractors");
    if (capacity <= 0 || capacity >= 65536 || !d_keys_l || !d_desc_l || !d_n_l || !d_keys_r || !d_desc_r || !d_n_r || !d_u_right || !d_depth)
        return fail(ORBFE_ERR_INVALID, "bad arguments (capacity must be < 65536)");
    cudaStream_t st = stream ? (cudaStream_t)stream : left->sCompute;
    const size_t nsad = (size_t)B * capacity;                 // SAD per left keypoint, then the candidate index
    const size_t need = nsad + orbfe_stereo_index_ints(left->g, B, capacity);
    if (need > left->stereoSadElems) {
        CK(cudaStreamSynchronize(st));
        if (left->d_stereoSad) cudaFree(left->d_stereoSad);
        left->d_stereoSad = nullptr; left->stereoSadElems = 0;
        CK(cudaMalloc(&left->d_stereoSad, need * sizeof(int)));
        left->stereoSadElems = need;
    }
    orbfe_launch_stereo_batch(left->g, (left->lastBufs ? left->lastBufs : &left->bufs)->pyr,
                              (right->lastBufs ? right->lastBufs : &right->bufs)->pyr, B, d_keys_l, (const uint32_t*)d_desc_l, d_n_l, d_keys_r,
                              (const uint32_t*)d_desc_r, d_n_r, capacity, mbf, mb, d_u_right, d_depth, left->d_stereoSad,
                              left->d_stereoSad + nsad, st);
    left->launches += 3;
    CK(cudaGetLastError());
    return ORBFE_OK;
}

int orbfe_frame_geometry(const OrbfeExtractor* h, int* cells, int* slots, int* kpcap, unsigned long long* pyr_stride,
                         unsigned long long* per_frame_bytes) {
    if (!h || !h->haveGeom) return fail(ORBFE_ERR_INVALID, "no geometry yet");
    if (cells) *cells = h->g.cellsPerFrame;
    if (slots) *slots = (int)h->g.slotsPerFrame;
    if (kpcap) *kpcap = h->g.kpCapFrame;
    if (pyr_stride) *pyr_stride = h->g.pyrStride;
    if (per_frame_bytes) *per_frame_bytes = h->perFrameBytes;
    return ORBFE_OK;
}

}  // extern "C"
