// octree.cu -- ORBextractor::DistributeOctTree (/root/reference/src/ORBextractor.cc:711-1057) and
// the keypoint bookkeeping around it (:1167-1201), one CTA per (pyramid level, frame).
//
// The CTA first turns the per-cell candidate slots written by k_fast_cells into the level's
// candidate list in the reference's emission order (exclusive scan over the cell counts, then a
// coalesced gather), then runs the CTA-parallel quadtree of octree_core.h on it.  Node tables
// live in shared memory (or in a global scratch block when nfeatures is so large that they do
// not fit); the points and their labels stay in global memory (L2 resident).
#include <algorithm>
#include <cstdlib>

#include "octree_core.h"
#include "orbfe_internal.h"

namespace {

// CTA size is chosen at launch: 128 threads per tree give the best throughput on large batches, 256 halve the
// critical path of the level-0 tree when only a few frames are in flight (single-frame calls: 0.44 -> 0.40 ms).
constexpr int OC_THREADS = OC_MAX_NT;    // upper bound (launch bounds, scratch sizes)
constexpr int OC_THREADS_BATCH = 128;

__device__ __forceinline__ int block_exclusive_scan(int v, int* warpSums, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpSums[wid] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
        const int s = warpSums[w];
        if (w < wid) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

// kShared: the node tables live in dynamic shared memory (the pointer provenance stays visible to
// the compiler, so table accesses are LDS/STS/ATOMS instead of generic loads); otherwise in a
// global scratch block (very large nfeatures).
template <bool kShared>
__global__ void __launch_bounds__(OC_THREADS)
k_octree(const __grid_constant__ OrbfeFrameGeom g, const uint32_t* __restrict__ slots,
         const int* __restrict__ cellCount, uint32_t* __restrict__ cand, uint32_t* __restrict__ pnode,
         int* __restrict__ candCount, uint32_t* __restrict__ kp, int* __restrict__ kpCount,
         char* ocGlobal, size_t ocGlobalStride) {
    extern __shared__ __align__(16) char smem[];
    __shared__ int warpSums[OC_THREADS / 32];
    __shared__ int s_outn;
    // grid = (frame, level): CTAs are dispatched x-fastest, so every frame's level-0 tree (the longest, ~0.2 ms) starts
    // in the first wave and the launch ends on the short top-level trees instead of on a level-0 straggler
    const int frame = blockIdx.x, level = blockIdx.y;
    const OrbfeLevelGeom& L = g.lv[level];
    const size_t fs = (size_t)frame * g.slotsPerFrame + L.slotBase;
    const uint32_t* cslots = slots + fs;
    uint32_t* pk = cand + fs;
    uint32_t* pn = pnode + fs;
    const int* cc = cellCount + (size_t)frame * g.cellsPerFrame + L.cellBase;
    const int nCells = L.nCols * L.nRows;

    // ---- candidate list in emission order: scan the cell counts, gather the slots ----
    int running = 0;
    for (int base = 0; base < nCells; base += blockDim.x) {
        const int c = base + threadIdx.x;
        const int v = c < nCells ? cc[c] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, warpSums, &tot);
        if (c < nCells) pn[c] = (uint32_t)(running + ex);  // pnode doubles as offset scratch
        running += tot;
    }
    const int n = running;
    __syncthreads();
    {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int c = wid; c < nCells; c += (int)(blockDim.x >> 5)) {
            const int cnt = cc[c];
            const int off = (int)pn[c];
            const uint32_t* s = cslots + (size_t)c * L.cellCap;
            for (int k = lane; k < cnt; k += 32) pk[off + k] = s[k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) candCount[frame * g.nlevels + level] = n;
    if (n == 0) {
        if (threadIdx.x == 0) kpCount[frame * g.nlevels + level] = 0;
        return;
    }

    // ---- quadtree ----
    const int M = L.ocM;
    char* mem = kShared ? smem : ocGlobal + ((size_t)frame * g.nlevels + level) * ocGlobalStride;
    OcWork w;
    oc_carve(w, mem, M);
    int* out_idx = w.cc;     // the child tables are dead once the list is final: reuse them for the
    int* best = w.cpos;      // per-node winner selection
    w.pk = pk;
    w.pnode = pn;
    w.n = n;
    oc_distribute(w, L.maxBX - ORBFE_FAST_BORDER, L.maxBY - ORBFE_FAST_BORDER, L.nIni, L.hX, L.nfeat,
                  out_idx, &s_outn, best);
    const int outn = s_outn;
    uint32_t* kpo = kp + (size_t)frame * g.kpCapFrame + L.kpBase;
    for (int k = threadIdx.x; k < outn && k < L.kpCap; k += blockDim.x) kpo[k] = pk[out_idx[k]];
    if (threadIdx.x == 0) kpCount[frame * g.nlevels + level] = min(outn, L.kpCap);
}

// Stand-alone DistributeOctTree on caller-supplied candidates (orbfe_debug_octree).
__global__ void __launch_bounds__(OC_THREADS)
k_octree_debug(const uint32_t* pk, uint32_t* pnode, int n, int width, int height, int nIni, float hX,
               int N, int M, int* out, int* outn, char* tables) {
    OcWork w;
    oc_carve(w, tables, M);
    int* out_idx = w.cc;
    int* best = w.cpos;
    w.pk = pk;
    w.pnode = pnode;
    w.n = n;
    __shared__ int s_outn;
    oc_distribute(w, width, height, nIni, hX, N, out_idx, &s_outn, best);
    for (int k = threadIdx.x; k < s_outn; k += blockDim.x) out[k] = out_idx[k];
    if (threadIdx.x == 0) *outn = s_outn;
}

size_t oc_total_bytes(int M) { return oc_shared_bytes(M) + 64; }

}  // namespace

size_t orbfe_octree_table_bytes(int M) { return (oc_total_bytes(M) + 255) & ~(size_t)255; }

int orbfe_octree_prepare(OrbfeFrameGeom& g) {
    int Mmax = 1;
    for (int l = 0; l < g.nlevels; l++) Mmax = g.lv[l].ocM > Mmax ? g.lv[l].ocM : Mmax;
    g.ocMmax = Mmax;
    const size_t need = oc_total_bytes(Mmax);
    if (need <= 200 * 1024) {
        g.ocShared = (int)need;
        if (need > 48 * 1024 &&
            cudaFuncSetAttribute(k_octree<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return -1;
    } else {
        g.ocShared = 0;
    }
    return (int)need;
}

void orbfe_launch_octree(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                         long long* launches) {
    int nt = B >= 128 ? OC_THREADS_BATCH : OC_THREADS;
    if (const char* ev = getenv("ORBFE_OC_THREADS")) nt = std::min(OC_THREADS, std::max(32, atoi(ev) & ~31));   // tuning
    if (g.ocShared)
        k_octree<true><<<dim3(B, g.nlevels), nt, g.ocShared, st>>>(
            g, b.slots, b.cellCount, b.cand, b.pnode, b.candCount, b.kp, b.kpCount, b.ocGlobal, b.ocGlobalStride);
    else
        k_octree<false><<<dim3(B, g.nlevels), nt, 0, st>>>(
            g, b.slots, b.cellCount, b.cand, b.pnode, b.candCount, b.kp, b.kpCount, b.ocGlobal, b.ocGlobalStride);
    ++*launches;
}

void orbfe_launch_octree_debug(const uint32_t* d_pk, uint32_t* d_pnode, int n, int width, int height,
                               int nIni, float hX, int N, int M, int* d_out, int* d_outn, char* d_tables,
                               cudaStream_t st) {
    k_octree_debug<<<1, OC_THREADS, 0, st>>>(d_pk, d_pnode, n, width, height, nIni, hX, N, M, d_out, d_outn, d_tables);
}
