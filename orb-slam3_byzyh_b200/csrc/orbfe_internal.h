// orbfe_internal.h -- geometry shared by the host planner (orbfe_api.cu) and the kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "../../include/orbfe.h"

#define ORBFE_XOFF 32  // column of the ROI origin inside a padded row (keeps ROI rows 16B aligned)
#define ORBFE_YOFF 19  // row of the ROI origin (EDGE_THRESHOLD)
#define ORBFE_FAST_BORDER 16  // minBorderX/Y = EDGE_THRESHOLD-3, ORBextractor.cc:1076-1079
#define ORBFE_HALF_PATCH 15   // HALF_PATCH_SIZE, ORBextractor.cc:77

// Per-level constants, all derived on the host exactly as the reference derives them
// (src/ORBextractor.cc:468-571 ctor, :1076-1095 FAST grid, :718-754 octree roots, :1692 sizes).
struct OrbfeLevelGeom {
    int w, h;            // ROI size
    int pitch;           // bytes per padded row (multiple of 16)
    unsigned off;        // byte offset of the padded level inside one frame's pyramid slab
    int nCols, nRows, wCell, hCell;  // FAST cell grid
    int maxBX, maxBY;    // w-16, h-16
    int cellBase;        // first cell id of this level (frame-wide numbering)
    int cellCap;         // candidate slots per cell = ceil(wCell/2)*ceil(hCell/2)
    unsigned slotBase;   // first candidate slot of this level (frame-wide numbering)
    int candCap;         // nCols*nRows*cellCap
    int nfeat;           // mnFeaturesPerLevel[level]
    int nIni;            // octree roots
    float hX;
    int ocM;             // octree node capacity
    int kpBase, kpCap;   // retained-keypoint slots (frame-wide numbering)
    float scale;         // mvScaleFactor[level]
    float invScale;      // mvInvScaleFactor[level]
    float kpsize;        // (float)(int)(PATCH_SIZE*mvScaleFactor[level])
    unsigned xtab, ytab; // offsets (in entries) of the resize tables of this level
    int mode;            // resize path: 0 = bilinear taps, 1 = exact 2x2 area, 2 = identity copy
    int fastTaps;        // mode 0: every aligned group of 4 destination columns taps <= 8 adjacent source bytes
    // tiled resize (k_resize_tile): per 128-column / 32-row block of the padded destination the bounding box of its
    // source taps (entries of the tap table: s = first padded source column, 16-aligned / first source ROW of the
    // block, s1 = last), and the TMA box that covers the largest of them; rzBoxW == 0: the level uses the older kernels
    unsigned rzXblk, rzYblk;
    int rzBoxW, rzBoxH;
    int blurTileBase, blurTilesX, blurTilesY;   // tile numbering of the blur kernel
};

struct OrbfeFrameGeom {
    int nlevels, rows, cols;
    int iniTh, minTh;
    unsigned long long pyrStride;   // bytes per frame in the pyramid / blurred slabs
    int cellsPerFrame;
    unsigned slotsPerFrame;
    int kpCapFrame;
    int blurTiles;
    int ocShared;                   // dynamic shared bytes of the octree kernel (0 = tables in global)
    int ocMmax;
    OrbfeLevelGeom lv[ORBFE_MAX_LEVELS];
};

// One bilinear tap table entry (per destination column or row): source indices and the two
// 11-bit weights, computed on the host in double precision exactly like OpenCV's resize.
struct OrbfeTap {
    short s;   // source index (clamped)
    short a0;  // weight of src[s]
    short a1;  // weight of src[s1]
    short s1;  // second source index (== s when clamped)
};

// Work item of the describe kernel: which level pixel to describe and where to put it.
struct OrbfeWork {
    short x, y;        // level coordinates (ROI)
    short level;
    short pad;
    int dst;           // index into the frame's output slab
    float angle;       // degrees
};

#define ORBFE_RZ_DW 128    // destination columns per CTA of the tiled resize (32 lanes x 4)
#define ORBFE_RZ_DH 32     // destination rows per CTA
#define ORBFE_BLUR_TW 120  // output columns per warp of the blur kernel
#define ORBFE_BLUR_TH 32   // output rows per warp strip

// Device buffers of one chunk of frames (all frame-major).
// TMA descriptors of the padded pyramid levels of one chunk buffer set: level l as a 3-D byte tensor
// (padded columns, padded rows, frames); k_fast_cells fetches a whole cell (+ ring halo) with one instruction.
struct OrbfeFastMaps {
    CUtensorMap m[ORBFE_MAX_LEVELS];
};

struct OrbfeChunkBufs {
    OrbfeFastMaps fastMaps;   // per level: box = one FAST cell + halo (fast.cu)
    OrbfeFastMaps resizeMaps; // m[l]: level l as the SOURCE of level l+1's tiled resize, box = rzBoxW x rzBoxH of level l+1
    uint8_t* pyr;        // [B][pyrStride]  padded pyramid levels
    uint8_t* blur;       // [B][pyrStride]  blurred levels (same geometry, ROI only)
    uint32_t* slots;     // [B][slotsPerFrame] per-cell candidate slots (packed)
    int* cellCount;      // [B][cellsPerFrame]
    uint32_t* cand;      // [B][slotsPerFrame] per-level compacted candidates (emission order)
    uint32_t* pnode;     // [B][slotsPerFrame] octree point labels
    int* candCount;      // [B][nlevels]
    uint32_t* kp;        // [B][kpCapFrame] retained candidates (packed), list order
    int* kpCount;        // [B][nlevels]
    OrbfeWork* work;     // [B][kpCapFrame]
    char* ocGlobal;      // octree tables when they do not fit in shared memory (else null)
    size_t ocGlobalStride;
};

// Records the thread's last error string (orbfe_last_error) and returns `code`.
int orbfe_fail(int code, const char* what, cudaError_t e);

// knn_umma.cu: the kNN-2 partial pass on the tensor cores.  _parts: (key0, key1) pairs per query the kernel writes for this
// size, 0 = size left to the scalar kernel; _enqueue: returns that number, or a negative ORBFE error code.
int orbfe_knn2_umma_parts(int nq, int nt);
int orbfe_knn2_umma_enqueue(const uint32_t* d_query, int nq, const uint32_t* d_train, int nt, uint32_t* d_partial, cudaStream_t st);
int orbfe_knn2_umma_batch_enqueue(const uint32_t* d_desc_q, const int* d_q_begin, const int* d_q_end, const uint32_t* d_desc_t,
                                  const int* d_t_begin, const int* d_t_end, int B, int capacity, int32_t* d_idx2, int32_t* d_dist2,
                                  int32_t* d_match, cudaStream_t st);

// ---- kernel launchers (each enqueues on `st`, no synchronisation) ----------------------------
// Level 0 source of orbfe_launch_pyramid: the frames as given, or rectified on the fly through CV_32FC1 maps.
struct OrbfeRectify {
    const float *mapx = nullptr, *mapy = nullptr;   // rectified rows x cols (= the geometry's level 0), device
    int srcRows = 0, srcCols = 0;                   // size of the raw input frames
};
void orbfe_launch_pyramid(const OrbfeFrameGeom& g, const OrbfeTap* taps, const uint8_t* d_images,
                          size_t step, size_t frameStride, const OrbfeChunkBufs& b, int B,
                          cudaStream_t st, long long* launches, const OrbfeRectify* rect = nullptr);
// Fills b.fastMaps for the buffer set (needs the driver's cuTensorMapEncodeTiled, resolved at run time).
int orbfe_fast_make_maps(const OrbfeFrameGeom& g, OrbfeChunkBufs& b, int frames);
int orbfe_resize_make_maps(const OrbfeFrameGeom& g, OrbfeChunkBufs& b, int frames);
int orbfe_make_level_maps(const OrbfeFrameGeom& g, const uint8_t* pyr, int frames, const int* boxW, const int* boxH,
                          OrbfeFastMaps& maps);
// cv::FAST per cell + NMS + the minThFAST retry + ordered emission, one warp per cell (fast.cu)
struct OrbfeFastCell {   // one FAST cell of a frame (frame-wide numbering): interior in ROI coordinates, nbx == 0 = skipped
    uint16_t x0, y0;
    uint8_t nbx, nby, level, pad;
};
void orbfe_fast_cell_table(const OrbfeFrameGeom& g, std::vector<OrbfeFastCell>& out);
void orbfe_launch_fast(const OrbfeFrameGeom& g, const OrbfeFastCell* cells, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                       long long* launches);
void orbfe_launch_octree(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                         long long* launches);
void orbfe_launch_blur(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                       long long* launches);
void orbfe_launch_layout(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, int lap0, int lap1,
                         OrbfeKeyPoint* d_kps, int capacity, int* d_n, int* d_mono, cudaStream_t st,
                         long long* launches);
void orbfe_launch_describe(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B,
                           OrbfeKeyPoint* d_kps, uint8_t* d_desc, int capacity, cudaStream_t st,
                           long long* launches);
int orbfe_octree_prepare(OrbfeFrameGeom& g);  // fills ocShared/ocMmax, sets the func attribute
size_t orbfe_octree_table_bytes(int M);
void orbfe_launch_octree_debug(const uint32_t* d_pk, uint32_t* d_pnode, int n, int width, int height,
                               int nIni, float hX, int N, int M, int* d_out, int* d_outn, char* d_tables,
                               cudaStream_t st);

// One ORBextractor instance (opaque to the C ABI).
struct OrbfeExtractor {
    int nfeatures, nlevels, iniTh, minTh, device;
    double scaleFactor;  // the reference stores the float ctor argument in a double member
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> nfeat;

    bool haveGeom = false;
    OrbfeFrameGeom g;
    OrbfeTap* d_taps = nullptr;
    OrbfeFastCell* d_cells = nullptr;   // per-cell FAST geometry of the current frame size
    int* d_stereoSad = nullptr;          // scratch of orbfe_stereo_match_batch_device (grow only)
    size_t stereoSadElems = 0;
    size_t perFrameBytes = 0;

    OrbfeChunkBufs bufs = {};
    int chunkCap = 0;       // frames the chunk buffers hold
    void* slab = nullptr;   // one allocation behind bufs
    // Second set of chunk intermediates + compute stream: multi-chunk host batches alternate between the two, so
    // that the kernels of chunk k+1 fill the SMs while chunk k drains its tails (octree, top pyramid levels).
    OrbfeChunkBufs bufs2 = {};
    int chunkCap2 = 0;
    void* slab2 = nullptr;
    cudaStream_t sCompute2 = nullptr;

    // staging of the host-pointer API (double buffered)
    uint8_t* d_in[2] = {nullptr, nullptr};
    size_t inBytes = 0;
    OrbfeKeyPoint* d_okps[2] = {nullptr, nullptr};
    uint8_t* d_odesc[2] = {nullptr, nullptr};
    int* d_on[2] = {nullptr, nullptr};
    int* d_omono[2] = {nullptr, nullptr};
    int outFrames = 0;
    size_t outElems = 0;

    cudaStream_t sCompute = nullptr, sH2D = nullptr, sD2H = nullptr;
    cudaEvent_t evIn[2] = {}, evInFree[2] = {}, evDone[2] = {}, evOutFree[2] = {};
    bool profiling = false;
    // per-kernel timing: a ring of event sets, one set per enqueued chunk while profiling is on
    static const int kProfSets = 64;
    cudaEvent_t evStage[kProfSets][ORBFE_NUM_STAGES + 1] = {};
    int profCount = 0;      // chunks recorded since profiling was switched on
    int lastFrames = 0;
    const OrbfeChunkBufs* lastBufs = nullptr;   // buffer set of the last enqueued chunk (bufs or bufs2)
    long long launches = 0;
    size_t maxBytes = (size_t)6 << 30;

    // Stereo rectification fused into pyramid level 0 (orbfe_extractor_set_rectification): CV_32FC1 maps of the
    // rectified size on the device; input images then have srcRows x srcCols of their own.
    float *d_mapx = nullptr, *d_mapy = nullptr;
    int rectRows = 0, rectCols = 0;
    int srcRows = 0, srcCols = 0;   // size of the frames of the current call

    // The per-frame call (Frame::ExtractORB: one or a few frames, host pointers) replays its 15 kernel launches as one
    // CUDA graph: the launch sequence only depends on the geometry, the lapping area and the staging buffers.
    struct GraphKey {
        int rows = 0, cols = 0, B = 0, lap0 = 0, lap1 = 0, capacity = 0;
        const void *in = nullptr, *kps = nullptr, *desc = nullptr, *slab = nullptr, *map = nullptr;
        bool operator==(const GraphKey& o) const {
            return rows == o.rows && cols == o.cols && B == o.B && lap0 == o.lap0 && lap1 == o.lap1 && capacity == o.capacity &&
                   in == o.in && kps == o.kps && desc == o.desc && slab == o.slab && map == o.map;
        }
    };
    GraphKey graphKey;
    cudaGraphExec_t graphExec = nullptr;
    long long graphLaunches = 0;

    // Host batches in flight (orbfe_extract_batch_submit / orbfe_extract_batch_wait): one completion event on sD2H per
    // submit, and what the wait has to check.  While batches are in flight the staging slots keep alternating across
    // submits (chunkSeq), so the H2D of the next batch runs under the kernels of the current one.
    static const int kMaxPending = 4;
    struct Pending { const int* n_out = nullptr; int B = 0, capacity = 0; };
    Pending pending[kMaxPending];
    cudaEvent_t evPending[kMaxPending] = {};
    int pendHead = 0, pendCount = 0;
    long long chunkSeq = 0;
};


// Frame::ComputeStereoMatches kernels (stereo.cu); pointers are device pointers.
size_t orbfe_stereo_index_ints(const OrbfeFrameGeom& g, int B, int capacity);   // scratch ints of the candidate index
void orbfe_launch_stereo(const OrbfeFrameGeom& g, const uint8_t* pyrL, const uint8_t* pyrR,
                         const OrbfeKeyPoint* keysL, const uint32_t* descL, int N, const OrbfeKeyPoint* keysR,
                         const uint32_t* descR, int Nr, float mbf, float mb, float* uRight, float* depth,
                         int* sad, int* idx, cudaStream_t st);
void orbfe_launch_stereo_batch(const OrbfeFrameGeom& g, const uint8_t* pyrL, const uint8_t* pyrR, int B,
                               const OrbfeKeyPoint* keysL, const uint32_t* descL, const int* nL, const OrbfeKeyPoint* keysR,
                               const uint32_t* descR, const int* nR, int capacity, float mbf, float mb, float* uRight,
                               float* depth, int* sad, int* idx, cudaStream_t st);
