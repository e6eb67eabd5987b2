// octree_core.h -- DistributeOctTree as a CTA-parallel algorithm.
//
// Reproduces, bit for bit (retained set AND list order), the reference's sequential quadtree
// culling: ExtractorNode::DivideNode (/root/reference/src/ORBextractor.cc:602-674),
// compareNodes (:676-697) and ORBextractor::DistributeOctTree (:711-1057), including
//   * std::list push_front ordering of children and in-place survival of bNoMore nodes,
//   * the per-pass nToExpand counter and the `size + 3*nToExpand > N` switch (:932),
//   * the phase-2 loop driven by an UNSTABLE std::sort: libstdc++'s introsort
//     (median-of-3, threshold 16, heapsort fallback, final insertion sort) is emulated
//     step for step (GCC 13 bits/stl_algo.h:1848-1951, bits/stl_heap.h) so equal-key nodes
//     land where the reference's binary puts them,
//   * first-wins max-response selection per node (:1028-1053).
//
// The sequential list algorithm is re-expressed as whole-list passes: every pass computes
// the four child populations of all expandable nodes with one data-parallel sweep over the
// points, then one thread rebuilds the (small) node list with the exact order the reference's
// push_front/erase sequence would produce, then all points are re-labelled in parallel.
//
// One CTA solves one (frame, level).  The same source compiles for the host (single
// "thread") so the logic is unit-tested on CPU against the reference; on the device
// OC_PAR_FOR strides over threadIdx.x and OC_SYNC is __syncthreads().
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OC_HD __host__ __device__ __forceinline__
#else
#define OC_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define OC_TID ((int)threadIdx.x)
#define OC_NT ((int)blockDim.x)
#define OC_SYNC() __syncthreads()
#define OC_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define OC_ATOMIC_ADD_RET(p, v) atomicAdd((p), (v))
#define OC_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define OC_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define OC_FMUL(a, b) __fmul_rn((a), (b))
#define OC_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define OC_TID 0
#define OC_NT 1
#define OC_SYNC() ((void)0)
#define OC_ATOMIC_ADD(p, v) (*(p) += (v))
#define OC_ATOMIC_ADD_RET(p, v) ((*(p) += (v)) - (v))
#define OC_ATOMIC_MAX(p, v) (*(p) = (*(p) > (v) ? *(p) : (v)))
#define OC_ATOMIC_MIN(p, v) (*(p) = (*(p) < (v) ? *(p) : (v)))
#define OC_FMUL(a, b) ((a) * (b))
#define OC_FDIV(a, b) ((a) / (b))
#endif
#define OC_PAR_FOR(i, n) for (int i = OC_TID; i < (n); i += OC_NT)
#ifndef OC_MARK   // tools/octree_phase_probe.cu defines it to accumulate clock64() per stage; nothing in the product
#define OC_MARK(slot) ((void)0)
#endif

// Packed FAST candidate: score[31:24] | y[23:12] | x[11:0] (window coordinates).
#define OC_PK_X(p) ((int)((p) & 0xFFFu))
#define OC_PK_Y(p) ((int)(((p) >> 12) & 0xFFFu))
#define OC_PK_S(p) ((int)((p) >> 24))
#define OC_PACK(x, y, s) (((uint32_t)(s) << 24) | ((uint32_t)(y) << 12) | (uint32_t)(x))

struct OcNodes {  // the two node-list buffers (structure of arrays); buffer b = entries [b*M, (b+1)*M)
    short* x0;
    short* x1;
    short* y0;
    short* y1;
    int* cnt;
};

struct OcWork {
    const uint32_t* pk;  // n packed candidates, emission order (index == tie-break key)
    uint32_t* pnode;     // n: node index [23:0] | quadrant [31:30]
    int n;
    int M;               // node capacity: max(N + 3, 4 * nIni) + 1
    OcNodes nd;          // double-buffered node list: buffer b lives at index offset b*M (no pointer tables
                         // indexed at run time, which would push this struct into local memory)
    int* cc;             // [4M] child populations of the current pass
    int* cpos;           // [4M] child -> index in the next list (-1 = empty)
    int* remap;          // [M]  undivided node -> index in the next list (-1 = divided)
    uint64_t* vs;        // [M]  expandable nodes, creation order: key<<32 | node
    uint64_t* vs2;       // [M]  scratch for the next creation-order list
    int* sc;             // [16] shared scalars
    int* part;           // [3 * OC_MAX_NT] per-thread partial sums of the parallel list rebuild
};
enum { OC_SIZE = 0, OC_CUR = 1, OC_NV = 2, OC_STATE = 3, OC_NTOEXP = 4, OC_P2_NDIV = 5, OC_P2_SIZE = 6, OC_P2_CHILDREN = 7, OC_SORT_CNT = 8 /* 8..10 */ };
#define OC_MAX_NT 256   // largest CTA size oc_distribute may be called with
enum { OC_ST_PHASE1 = 0, OC_ST_PHASE2 = 1, OC_ST_DONE = 2 };

static OC_HD size_t oc_shared_bytes(int M) {
    // 2 buffers x (4 shorts + 1 int) + cc + cpos (4 ints each) + remap + vs + vs2 + scalars
    return (size_t)M * (2 * (4 * sizeof(short) + sizeof(int)) + 8 * sizeof(int) + sizeof(int) +
                        2 * sizeof(uint64_t)) + 16 * sizeof(int) + 64 + 3 * OC_MAX_NT * sizeof(int);
}

// ---- libstdc++ std::sort emulation on key<<32|payload words (compare on the key only) ----
#define OC_LESS(a, b) (((a) >> 32) < ((b) >> 32))

static OC_HD void oc_unguarded_linear_insert(uint64_t* a, int last) {
    uint64_t val = a[last];
    int next = last - 1;
    while (OC_LESS(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}
static OC_HD void oc_insertion_sort(uint64_t* a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (OC_LESS(a[i], a[first])) {
            uint64_t val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];
            a[first] = val;
        } else {
            oc_unguarded_linear_insert(a, i);
        }
    }
}
static OC_HD void oc_adjust_heap(uint64_t* a /*first*/, int hole, int len, uint64_t value) {
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (OC_LESS(a[second], a[second - 1])) second--;
        a[hole] = a[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        a[hole] = a[second - 1];
        hole = second - 1;
    }
    int parent = (hole - 1) / 2;  // __push_heap
    while (hole > top && OC_LESS(a[parent], value)) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = value;
}
static OC_HD void oc_heap_sort(uint64_t* a, int n) {  // __partial_sort(first, last, last)
    if (n >= 2) {
        int parent = (n - 2) / 2;
        while (true) {
            uint64_t v = a[parent];
            oc_adjust_heap(a, parent, n, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = n;
    while (last > 1) {
        --last;
        uint64_t v = a[last];
        a[last] = a[0];
        oc_adjust_heap(a, 0, last, v);
    }
}
static OC_HD void oc_swap(uint64_t* a, int i, int j) {
    uint64_t t = a[i];
    a[i] = a[j];
    a[j] = t;
}
// std::__introsort_loop: after it every element sits inside its final partition of at most 16 elements (or in a
// heap-sorted range).  The pivot key is read once per partition (the partition never moves a[first]).
static OC_HD void oc_introsort_loop(uint64_t* a, int n) {
    if (n <= 0) return;
    int lg = 0;
    while ((n >> (lg + 1)) != 0) lg++;
    // explicit stack instead of the recursion on the right-hand partition
    int stF[64], stL[64], stD[64];
    int sp = 0;
    stF[0] = 0; stL[0] = n; stD[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stF[sp], last = stL[sp], depth = stD[sp];
        while (last - first > 16) {
            if (depth == 0) {
                oc_heap_sort(a + first, last - first);
                break;
            }
            --depth;
            // __unguarded_partition_pivot
            const int mid = first + (last - first) / 2;
            {   // __move_median_to_first(first, first+1, mid, last-1)
                const int A = first + 1, B = mid, C = last - 1;
                if (OC_LESS(a[A], a[B])) {
                    if (OC_LESS(a[B], a[C])) oc_swap(a, first, B);
                    else if (OC_LESS(a[A], a[C])) oc_swap(a, first, C);
                    else oc_swap(a, first, A);
                } else if (OC_LESS(a[A], a[C])) oc_swap(a, first, A);
                else if (OC_LESS(a[B], a[C])) oc_swap(a, first, C);
                else oc_swap(a, first, B);
            }
            const uint64_t pivot = a[first];
            int lo = first + 1, hi = last;
            while (true) {  // __unguarded_partition(first+1, last, pivot = first)
                while (OC_LESS(a[lo], pivot)) ++lo;
                --hi;
                while (OC_LESS(pivot, a[hi])) --hi;
                if (!(lo < hi)) break;
                oc_swap(a, lo, hi);
                ++lo;
            }
            const int cut = lo;
            if (sp < 64) { stF[sp] = cut; stL[sp] = last; stD[sp] = depth; ++sp; }
            last = cut;
        }
    }
}
static OC_HD void oc_std_sort(uint64_t* a, int n) {
    if (n <= 0) return;
    oc_introsort_loop(a, n);
    // __final_insertion_sort
    if (n > 16) {
        oc_insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) oc_unguarded_linear_insert(a, i);
    } else {
        oc_insertion_sort(a, 0, n);
    }
}

// Exclusive prefix sums (thread order) of three per-thread counters over the CTA; every thread gets the totals.
// Device: warp shuffles + one shared word per warp and counter (w.part), two barriers.  Host build: one "thread".
static OC_HD void oc_cta_exscan3(int* part, int& a, int& b, int& c, int& ta, int& tb, int& tc) {
#if defined(__CUDA_ARCH__)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (int)((blockDim.x + 31) >> 5);
    int ia = a, ib = b, ic = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int xa = __shfl_up_sync(0xffffffffu, ia, o), xb = __shfl_up_sync(0xffffffffu, ib, o),
                  xc = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) { ia += xa; ib += xb; ic += xc; }
    }
    if (lane == 31) { part[wid] = ia; part[32 + wid] = ib; part[64 + wid] = ic; }
    __syncthreads();
    int pa = 0, pb = 0, pc = 0;
    ta = tb = tc = 0;
    for (int k = 0; k < nw; k++) {
        const int xa = part[k], xb = part[32 + k], xc = part[64 + k];
        if (k < wid) { pa += xa; pb += xb; pc += xc; }
        ta += xa; tb += xb; tc += xc;
    }
    a = pa + ia - a; b = pb + ib - b; c = pc + ic - c;
    __syncthreads();   // part is free again
#else
    (void)part;
    ta = a; tb = b; tc = c;
    a = b = c = 0;
#endif
}

// ---- std::__introsort_loop by the whole CTA ------------------------------------------------
// The partitions of one recursion level are disjoint ranges, so they are independent: one GROUP (a warp; the single
// host "thread") takes one range, and the levels are separated by a CTA barrier.  Inside a range the Hoare partition
// of __unguarded_partition is order-free too: its k-th swap exchanges the k-th element from the left that is not
// below the pivot (L[k]) with the k-th element from the right that is not above it (R[k]), for as long as
// L[k] < R[k] -- earlier swaps only touch positions outside (L[k], R[k]).  So a group ranks both sets with ballots,
// counts the swaps K (the condition holds on a prefix: L ascends, R descends), swaps the K pairs at once and takes
// cut = min(L[K], R[K-1]): the first position after the last swap whose CURRENT value is not below the pivot.
#if defined(__CUDA_ARCH__)
#define OC_LANE ((int)(threadIdx.x & 31))
#define OC_GS 32
#define OC_GID ((int)(threadIdx.x >> 5))
#define OC_NG ((int)(blockDim.x >> 5))
#define OC_GSYNC() __syncwarp()
#else
#define OC_LANE 0
#define OC_GS 1
#define OC_GID 0
#define OC_NG 1
#define OC_GSYNC() ((void)0)
#endif
static OC_HD void oc_group_rank(bool f, int& rank, int& total) {
#if defined(__CUDA_ARCH__)
    const unsigned m = __ballot_sync(0xffffffffu, f);
    rank = __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
    total = __popc(m);
#else
    rank = 0;
    total = f ? 1 : 0;
#endif
}
// __unguarded_partition_pivot(first, last) by one group; Lp / Rp: int scratch indexed like a.  Returns the cut.
static OC_HD int oc_partition_group(uint64_t* a, int* Lp, int* Rp, int first, int last) {
    const int lane = OC_LANE;
    if (lane == 0) {   // __move_median_to_first(first, first+1, mid, last-1)
        const int A = first + 1, B = first + (last - first) / 2, C = last - 1;
        if (OC_LESS(a[A], a[B])) {
            if (OC_LESS(a[B], a[C])) oc_swap(a, first, B);
            else if (OC_LESS(a[A], a[C])) oc_swap(a, first, C);
            else oc_swap(a, first, A);
        } else if (OC_LESS(a[A], a[C])) oc_swap(a, first, A);
        else if (OC_LESS(a[B], a[C])) oc_swap(a, first, C);
        else oc_swap(a, first, B);
    }
    OC_GSYNC();
    const uint32_t pivot = (uint32_t)(a[first] >> 32);
    const int s = first + 1, len = last - s;
    int nL = 0, nR = 0;
    for (int c = 0; c < len; c += OC_GS) {
        const int i = c + lane;
        const int pl = s + i, pr = last - 1 - i;
        bool ge = false, le = false;
        if (i < len) {
            ge = (uint32_t)(a[pl] >> 32) >= pivot;
            le = (uint32_t)(a[pr] >> 32) <= pivot;
        }
        int r, t;
        oc_group_rank(ge, r, t);
        if (ge) Lp[s + nL + r] = pl;
        nL += t;
        oc_group_rank(le, r, t);
        if (le) Rp[s + nR + r] = pr;
        nR += t;
    }
    OC_GSYNC();
    const int m = nL < nR ? nL : nR;
    int K = 0;
    for (int c = 0; c < m; c += OC_GS) {
        const int k = c + lane;
        int r, t;
        oc_group_rank(k < m && Lp[s + k] < Rp[s + k], r, t);
        K += t;
    }
    int cut = last;   // the median guarantees an element >= pivot inside the range
    if (K < nL) cut = Lp[s + K];
    if (K > 0 && Rp[s + K - 1] < cut) cut = Rp[s + K - 1];
    for (int k = lane; k < K; k += OC_GS) oc_swap(a, Lp[s + k], Rp[s + k]);
    OC_GSYNC();
    return cut;
}

// std::sort(a, a+n) by the whole CTA, same permutation as libstdc++: the introsort loop runs level by level (above),
// and __final_insertion_sort is a plain insertion sort over the whole range, i.e. the STABLE sort of what the loop left
// behind -- every thread ranks one element (keys below it + equal keys before it).  The sorted sequence is returned in
// `out` (a is left partitioned, not sorted; `out` doubles as the partition scratch).  q: 4 * (n/16 + 2) ints of
// scratch for the two range queues, cnt: 3 shared counters.  All threads must call; ends with a barrier.
static OC_HD void oc_std_sort_cta(uint64_t* a, uint64_t* out, int n, int* q, int* cnt) {
    if (n > 16) {
        const int qcap = n / 16 + 2;
        int* Lp = (int*)out;
        int* Rp = Lp + n;
        if (OC_TID == 0) {
            int lg = 0;
            while ((n >> (lg + 1)) != 0) lg++;
            q[0] = 0; q[1] = n | ((2 * lg) << 24);
            cnt[0] = 1; cnt[1] = 0; cnt[2] = 0;
        }
        OC_SYNC();
        for (int level = 0;; level++) {
            const int rd = level % 3, wr = (level + 1) % 3, zr = (level + 2) % 3;
            const int ns = cnt[rd];
            if (ns == 0) break;
            if (OC_TID == 0) cnt[zr] = 0;
            const int* qr = q + (level & 1) * 2 * qcap;
            int* qw = q + ((level + 1) & 1) * 2 * qcap;
            for (int sg = OC_GID; sg < ns; sg += OC_NG) {
                const int first = qr[2 * sg], last = qr[2 * sg + 1] & 0xFFFFFF, depth = qr[2 * sg + 1] >> 24;
                if (depth == 0) {   // __partial_sort(first, last, last)
                    if (OC_LANE == 0) oc_heap_sort(a + first, last - first);
                    continue;
                }
                const int cut = oc_partition_group(a, Lp, Rp, first, last);
                if (OC_LANE == 0) {
                    if (last - cut > 16) {
                        const int e = OC_ATOMIC_ADD_RET(&cnt[wr], 1);
                        qw[2 * e] = cut; qw[2 * e + 1] = last | ((depth - 1) << 24);
                    }
                    if (cut - first > 16) {
                        const int e = OC_ATOMIC_ADD_RET(&cnt[wr], 1);
                        qw[2 * e] = first; qw[2 * e + 1] = cut | ((depth - 1) << 24);
                    }
                }
            }
            OC_SYNC();
        }
    }
    OC_SYNC();
    OC_PAR_FOR(i, n) {
        const uint64_t v = a[i];
        const uint32_t key = (uint32_t)(v >> 32);
        int r = 0;
        for (int j = 0; j < n; j++) {
            const uint32_t kj = (uint32_t)(a[j] >> 32);
            r += (kj < key || (kj == key && j < i)) ? 1 : 0;
        }
        out[r] = v;
    }
    OC_SYNC();
}

// ---- DivideNode geometry (:608-637) -------------------------------------------------------
static OC_HD int oc_half(int lo, int hi) {  // ceil(static_cast<float>(hi-lo)/2), exact in int
    const int d = hi - lo;
    return d >= 0 ? (d + 1) / 2 : -((-d) / 2);
}
// `j` and `pos` are absolute table indices (buffer offset already added).
static OC_HD void oc_make_child(const OcNodes& nd, int j, int q, int pos, int count) {
    const int X0 = nd.x0[j], X1 = nd.x1[j], Y0 = nd.y0[j], Y1 = nd.y1[j];
    const int mx = X0 + oc_half(X0, X1), my = Y0 + oc_half(Y0, Y1);
    nd.x0[pos] = (short)((q & 1) ? mx : X0);
    nd.x1[pos] = (short)((q & 1) ? X1 : mx);
    nd.y0[pos] = (short)((q & 2) ? my : Y0);
    nd.y1[pos] = (short)((q & 2) ? Y1 : my);
    nd.cnt[pos] = count;
}
static OC_HD void oc_copy_node(const OcNodes& nd, int j, int pos) {
    nd.x0[pos] = nd.x0[j]; nd.x1[pos] = nd.x1[j];
    nd.y0[pos] = nd.y0[j]; nd.y1[pos] = nd.y1[j];
    nd.cnt[pos] = nd.cnt[j];
}

// The point sweeps issue OC_ILP independent loads per thread before using any of them: the points
// live in global memory (L2) and a dependent load chain per point would expose its latency.
#define OC_ILP 4

// One sweep over the points per pass.  After the node list was rebuilt every point takes the index its node (or its
// child) has in the new list, and -- unless the distribution is finished -- is counted straight away into the child
// populations of the NEXT pass: its quadrant inside the new node, if that node is still expandable.  (Two sweeps per
// pass, one to count and one to re-label, read and wrote the labels twice and exposed the L2 latency twice.)
static OC_HD void oc_relabel_count(const OcWork& w, int* win) {
    const OcNodes& nd = w.nd;
    const int so = w.sc[OC_CUR] * w.M;     // offset of the (new) current list buffer
    const int size = w.sc[OC_SIZE];
    const bool count = w.sc[OC_STATE] != OC_ST_DONE;
    if (count) OC_PAR_FOR(k, 4 * size) w.cc[k] = 0;
    else OC_PAR_FOR(k, size) win[k] = 0;
    OC_SYNC();
    for (int i0 = OC_TID; i0 < w.n; i0 += OC_ILP * OC_NT) {
        uint32_t lab[OC_ILP], p[OC_ILP];
#pragma unroll
        for (int u = 0; u < OC_ILP; u++) {
            const int i = i0 + u * OC_NT;
            lab[u] = i < w.n ? w.pnode[i] : 0u;
            p[u] = i < w.n ? w.pk[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < OC_ILP; u++) {
            const int i = i0 + u * OC_NT;
            if (i < w.n) {
                const int jo = (int)(lab[u] & 0xFFFFFFu);
                const int r = w.remap[jo];
                const int j = r >= 0 ? r : w.cpos[4 * jo + (int)(lab[u] >> 30)];
                uint32_t out = (uint32_t)j;
                if (count && nd.cnt[so + j] > 1) {
                    const int X0 = nd.x0[so + j], X1 = nd.x1[so + j], Y0 = nd.y0[so + j], Y1 = nd.y1[so + j];
                    const int mx = X0 + oc_half(X0, X1), my = Y0 + oc_half(Y0, Y1);
                    const int q = (OC_PK_X(p[u]) < mx ? 0 : 1) + (OC_PK_Y(p[u]) < my ? 0 : 2);
                    out |= (uint32_t)q << 30;
                    OC_ATOMIC_ADD(&w.cc[4 * j + q], 1);
                }
                // last pass: the node's winner (:1028-1053) is the maximum of response << 24 | (2^24 - 1 - index), i.e.
                // the best response and, among equals, the first candidate
                if (!count) OC_ATOMIC_MAX((unsigned int*)&win[j], ((unsigned int)OC_PK_S(p[u]) << 24) | (unsigned int)(0xFFFFFF - i));
                w.pnode[i] = out;
            }
        }
    }
    OC_SYNC();
}

// Phase-1 list rebuild (:805-930), all threads.  The reference walks the list once, replacing
// every expandable node by its non-empty children (push_front) and leaving the others in place;
// the resulting order is: children of node size-1 (n4..n1), ..., children of node 0, then the
// undivided nodes in their old order.  Every position is therefore a prefix sum over the old list:
// three scans (non-empty children, undivided nodes, children with more than one point) give all
// threads their write offsets, so no thread walks the list alone.
static OC_HD void oc_rebuild_phase1(const OcWork& w, int N) {
    const int cur = w.sc[OC_CUR];
    const OcNodes& nd = w.nd;
    const int so = cur * w.M, dof = (cur ^ 1) * w.M;   // source / destination list buffers
    const int size = w.sc[OC_SIZE];
    const int nt = OC_NT, tid = OC_TID;
    const int chunk = (size + nt - 1) / nt;
    const int j0 = tid * chunk < size ? tid * chunk : size, j1 = j0 + chunk < size ? j0 + chunk : size;
    int sk = 0, sne = 0, sm = 0;
    for (int j = j0; j < j1; j++) {
        if (nd.cnt[so + j] > 1) {
            for (int q = 0; q < 4; q++) {
                const int c = w.cc[4 * j + q];
                sk += c > 0;
                sm += c > 1;
            }
        } else {
            sne++;
        }
    }
    int preK = sk, preNE = sne, preM = sm, totK, totNE, totM;
    oc_cta_exscan3(w.part, preK, preNE, preM, totK, totNE, totM);
    for (int j = j0; j < j1; j++) {
        if (nd.cnt[so + j] > 1) {
            int k = 0;
            for (int q = 0; q < 4; q++) k += w.cc[4 * j + q] > 0;
            preK += k;
            int pos = totK - preK;   // children of later nodes sit nearer the front (push_front)
            w.remap[j] = -1;
            for (int q = 3; q >= 0; q--) {
                const int c = w.cc[4 * j + q];
                if (c > 0) { oc_make_child(nd, so + j, q, dof + pos, c); w.cpos[4 * j + q] = pos++; }
                else w.cpos[4 * j + q] = -1;
            }
            for (int q = 0; q < 4; q++) {
                const int c = w.cc[4 * j + q];
                if (c > 1) {
                    const int cp = w.cpos[4 * j + q];
                    w.vs[preM++] = ((uint64_t)(((uint32_t)c << 13) | (uint32_t)nd.x0[dof + cp]) << 32) | (uint32_t)cp;
                }
            }
        } else {
            const int pos = totK + preNE++;
            oc_copy_node(nd, so + j, dof + pos);
            w.remap[j] = pos;
        }
    }
    OC_SYNC();
    if (tid == 0) {
        const int newSize = totK + totNE;
        int st = OC_ST_PHASE1;
        if (newSize >= N || newSize == size) st = OC_ST_DONE;
        else if (newSize + totM * 3 > N) st = OC_ST_PHASE2;   // :932, nToExpand == totM
        w.sc[OC_STATE] = st;
        w.sc[OC_NV] = totM;
        w.sc[OC_SIZE] = newSize;
        w.sc[OC_CUR] = cur ^ 1;
    }
}

// One phase-2 step (:934-1015): the expandable nodes are sorted by (population, x) and divided one at a time, largest
// first, until the list holds N nodes.  Whole CTA: the sort's final pass, the child tables of the divided nodes and the
// copies of the undivided ones are done by all threads; only the introsort partitioning and the short scan that finds
// how many nodes get divided run on one thread.
static OC_HD void oc_phase2(const OcWork& w, int N) {
    const OcNodes& nd = w.nd;
    const int cur = w.sc[OC_CUR];
    const int so = cur * w.M, dof = (cur ^ 1) * w.M;
    const int size = w.sc[OC_SIZE];
    const int np = w.sc[OC_NV];
    oc_std_sort_cta(w.vs, w.vs2, np, w.cpos, w.sc + OC_SORT_CNT);   // sorted sequence in vs2; vs and cpos are scratch
    OC_MARK(6);   // phase-2: sort
    int* ne = (int*)w.vs;               // [np] non-empty children of the t-th division (t = 0: largest key)
    int* off = ne + w.M;                // [np] first list position of its children
    OC_PAR_FOR(j, size) w.remap[j] = 0;
    OC_PAR_FOR(t, np) {
        const int j = (int)(w.vs2[np - 1 - t] & 0xFFFFFFFFu);
        int c = 0;
        for (int q = 0; q < 4; q++) c += (w.cc[4 * j + q] > 0);
        ne[t] = c;
    }
    OC_SYNC();
    if (OC_TID == 0) {
        int newSize = size, ndiv = 0;
        for (int t = 0; t < np; t++) {
            newSize += ne[t] - 1;
            ndiv++;
            if (newSize >= N) break;
        }
        // the last node divided owns the front of the list
        int pos = 0;
        for (int t = ndiv - 1; t >= 0; t--) { off[t] = pos; pos += ne[t]; }
        w.sc[OC_P2_NDIV] = ndiv;
        w.sc[OC_P2_SIZE] = newSize;
        w.sc[OC_P2_CHILDREN] = pos;
    }
    OC_SYNC();
    const int ndiv = w.sc[OC_P2_NDIV], newSize = w.sc[OC_P2_SIZE], nChildren = w.sc[OC_P2_CHILDREN];
    OC_PAR_FOR(t, ndiv) {
        const int j = (int)(w.vs2[np - 1 - t] & 0xFFFFFFFFu);
        w.remap[j] = -1;
        int pos = off[t];
        for (int q = 3; q >= 0; q--) {
            const int c = w.cc[4 * j + q];
            if (c > 0) { oc_make_child(nd, so + j, q, dof + pos, c); w.cpos[4 * j + q] = pos++; }
            else w.cpos[4 * j + q] = -1;
        }
    }
    OC_SYNC();
    OC_MARK(7);   // phase-2: divisions
    {   // undivided nodes keep their order behind the children
        const int nt = OC_NT, tid = OC_TID;
        const int chunk = (size + nt - 1) / nt;
        const int j0 = tid * chunk < size ? tid * chunk : size, j1 = j0 + chunk < size ? j0 + chunk : size;
        int pre = 0, d1 = 0, d2 = 0, t0, t1, t2;
        for (int j = j0; j < j1; j++) pre += (w.remap[j] != -1);
        oc_cta_exscan3(w.part, pre, d1, d2, t0, t1, t2);
        int pos = nChildren + pre;
        for (int j = j0; j < j1; j++)
            if (w.remap[j] != -1) { oc_copy_node(nd, so + j, dof + pos); w.remap[j] = pos++; }
    }
    const bool done = newSize >= N || newSize == size;
    if (OC_TID == 0) {
        int nv = 0;
        if (!done) {
            // creation-order list of the next step (rare: one step almost always reaches N); ne/off are dead
            for (int t = 0; t < ndiv; t++) {
                const int j = (int)(w.vs2[np - 1 - t] & 0xFFFFFFFFu);
                for (int q = 0; q < 4; q++) {
                    const int c = w.cc[4 * j + q];
                    if (c > 1) {
                        const int cp = w.cpos[4 * j + q];
                        w.vs[nv++] = ((uint64_t)(((uint32_t)c << 13) | (uint32_t)nd.x0[dof + cp]) << 32) | (uint32_t)cp;
                    }
                }
            }
        } else {
            w.sc[OC_STATE] = OC_ST_DONE;
        }
        w.sc[OC_NV] = nv;
        w.sc[OC_SIZE] = newSize;
        w.sc[OC_CUR] = cur ^ 1;
        OC_MARK(3);   // phase-2: copies + new list
    }
}

// Runs the whole distribution.  Output: out_idx[k] = candidate index retained by the k-th
// node of the final list (front to back); returns the number of nodes via *out_n (thread 0
// writes it; visible to all threads after the final OC_SYNC).
//   width/height = maxX-minX / maxY-minY of the level window, N = mnFeaturesPerLevel[level].
static OC_HD void oc_distribute(const OcWork& w, int width, int height, int nIni, float hX, int N,
                                int* out_idx, int* out_n, int* /*scratch, unused*/) {
    // ---- roots (:718-790) ----
    if (OC_TID == 0) {
        const OcNodes& nd = w.nd;
        for (int i = 0; i < nIni; i++) {
            nd.x0[i] = (short)(int)OC_FMUL(hX, (float)i);
            nd.x1[i] = (short)(int)OC_FMUL(hX, (float)(i + 1));
            nd.y0[i] = 0;
            nd.y1[i] = (short)height;
            nd.cnt[i] = 0;
        }
        w.sc[OC_CUR] = 0;
    }
    OC_SYNC();
    OC_PAR_FOR(i, w.n) {
        int r = (int)OC_FDIV((float)OC_PK_X(w.pk[i]), hX);
        w.pnode[i] = (uint32_t)r;
        OC_ATOMIC_ADD(&w.nd.cnt[r], 1);
    }
    OC_SYNC();
    if (OC_TID == 0) {  // erase empty roots, keep order
        const OcNodes& nd = w.nd;
        int pos = 0;
        for (int i = 0; i < nIni; i++) {
            w.cpos[4 * i] = w.cpos[4 * i + 1] = w.cpos[4 * i + 2] = w.cpos[4 * i + 3] = -1;
            if (nd.cnt[i] > 0) { oc_copy_node(nd, i, w.M + pos); w.remap[i] = pos++; }
            else w.remap[i] = -1;
        }
        w.sc[OC_CUR] = 1;
        w.sc[OC_SIZE] = pos;
        w.sc[OC_STATE] = OC_ST_PHASE1;
        w.sc[OC_NV] = 0;
    }
    OC_SYNC();
    oc_relabel_count(w, out_idx);
    OC_MARK(0);   // roots

    // ---- main loop (:805-1020) ----
    while (true) {
        const int state = w.sc[OC_STATE];
        if (state == OC_ST_DONE) break;   // (thread 0 rewrites the state only behind a barrier inside the rebuild)
        if (state == OC_ST_PHASE1) {
            oc_rebuild_phase1(w, N);
            OC_MARK(2);   // phase-1 rebuild (thread 0's share)
        } else {
            oc_phase2(w, N);
        }
        OC_SYNC();
        oc_relabel_count(w, out_idx);
        OC_MARK(4);   // relabel + child populations of the next pass
    }

    // ---- best response per node, first candidate wins ties (:1028-1053): selected by the last sweep ----
    const int size = w.sc[OC_SIZE];
    OC_PAR_FOR(k, size) out_idx[k] = 0xFFFFFF - (out_idx[k] & 0xFFFFFF);
    if (OC_TID == 0) *out_n = size;
    OC_SYNC();
    OC_MARK(5);   // winners
}

// Carve the node tables out of one contiguous (shared-memory) block of oc_shared_bytes(M).
static OC_HD void oc_carve(OcWork& w, void* mem, int M) {
    char* p = (char*)mem;
    w.M = M;
    w.vs = (uint64_t*)p; p += sizeof(uint64_t) * M;
    w.vs2 = (uint64_t*)p; p += sizeof(uint64_t) * M;
    w.cc = (int*)p; p += sizeof(int) * 4 * M;
    w.cpos = (int*)p; p += sizeof(int) * 4 * M;
    w.remap = (int*)p; p += sizeof(int) * M;
    w.nd.cnt = (int*)p; p += sizeof(int) * 2 * M;
    w.sc = (int*)p; p += sizeof(int) * 16;
    w.part = (int*)p; p += sizeof(int) * 3 * OC_MAX_NT;
    w.nd.x0 = (short*)p; p += sizeof(short) * 2 * M;
    w.nd.x1 = (short*)p; p += sizeof(short) * 2 * M;
    w.nd.y0 = (short*)p; p += sizeof(short) * 2 * M;
    w.nd.y1 = (short*)p; p += sizeof(short) * 2 * M;
}
