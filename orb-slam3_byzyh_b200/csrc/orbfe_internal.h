// orbfe_internal.h -- geometry shared by the host planner (orbfe_api.cu) and the kernels.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/orbfe.h"

#define ORBFE_XOFF 32  // column of the ROI origin inside a padded row (keeps ROI rows 16B aligned)
#define ORBFE_YOFF 19  // row of the ROI origin (EDGE_THRESHOLD)
#define ORBFE_FAST_BORDER 16  // minBorderX/Y = EDGE_THRESHOLD-3, ORBextractor.cc:1076-1079

// Per-level constants, all derived on the host exactly as the reference derives them
// (src/ORBextractor.cc:468-571 ctor, :1076-1095 FAST grid, :718-754 octree roots, :1692 sizes).
struct OrbfeLevelGeom {
    int w, h;            // ROI size
    int pitch;           // bytes per padded row
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
    float kpsize;        // (float)(int)(PATCH_SIZE*mvScaleFactor[level])
    unsigned xtab, ytab; // offsets (in entries) of the resize tables of this level
    int areaFast;        // exact 2x decimation -> OpenCV's area path
};

struct OrbfeFrameGeom {
    int nlevels, rows, cols;
    int iniTh, minTh;
    unsigned long long pyrStride;   // bytes per frame in the pyramid / blurred slabs
    int cellsPerFrame;
    unsigned slotsPerFrame;
    int kpCapFrame;
    OrbfeLevelGeom lv[ORBFE_MAX_LEVELS];
};

// One bilinear tap table entry (per destination column or row): source index and the two
// 11-bit weights, computed on the host in double precision exactly like OpenCV's resize.
struct OrbfeTap {
    short s;   // source index (clamped)
    short a0;  // weight of src[s]
    short a1;  // weight of src[s+1]
    short s1;  // second source index (== s when clamped)
};

// Work item of the describe kernel: which level pixel to describe and where to put it.
struct OrbfeWork {
    short x, y;        // level coordinates (ROI)
    short level;
    short pad;
    int dst;           // index into the frame's output slab
};
