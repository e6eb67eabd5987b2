// bow.cu -- bag-of-words path and its C ABI (include/orbfe.h):
//   Frame::ComputeBoW / KeyFrame::ComputeBoW        /root/reference/src/Frame.cc:984-998, src/KeyFrame.cc:101-111
//   DBoW2::TemplatedVocabulary::transform           Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1125-1197, 1226-1258
//   DBoW2::FORB::distance                           Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
//   ORBmatcher::SearchByBoW x2                      src/ORBmatcher.cc:260-494, 893-1044
//
// The vocabulary tree lives in HBM as flat arrays (descriptors 32 B per node, children in CSR form in the order
// DBoW2 pushed them, word id and weight per node; ORBvoc.txt: 1.1 M nodes = 35 MB + 13 MB).  transform() is a
// root-to-leaf descent with a Hamming argmin over the <= k children per level: one warp per feature, one lane
// per child, the key (distance << 20 | child position) makes the warp minimum DBoW2's first-smallest child.
// SearchByBoW joins two feature vectors node by node; inside a node the reference's loop is sequential (a frame
// feature taken by an earlier keyframe feature is skipped), across nodes it is independent because every feature
// sits in exactly one node: one warp per common node, keyframe features in order, lanes over the frame features.
#include <algorithm>
#include <vector>

#include "kb8_core.h"
#include "orbfe_internal.h"
#include "scratch.h"

struct OrbfeVocabulary {
    int device = 0, k = 0, L = 0, nNodes = 0, maxChildren = 0;
    uint32_t* desc = nullptr;    // nNodes x 8 words
    int* childStart = nullptr;   // nNodes + 1
    int* children = nullptr;     // nNodes - 1
    int* wordId = nullptr;       // -1 for inner nodes
    double* weight = nullptr;
};

namespace {

constexpr int HISTO = 30;

int bfail(int code, const char* what, cudaError_t e = cudaSuccess) { return orbfe_fail(code, what, e); }
#define BCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return bfail(ORBFE_ERR_CUDA, #call, e_);  \
    } while (0)

__device__ __forceinline__ int hamming8w(const uint32_t* a, const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup), :1226-1258
__global__ void __launch_bounds__(128)
k_bow_transform(const uint32_t* __restrict__ vdesc, const int* __restrict__ childStart, const int* __restrict__ children,
                const int* __restrict__ wordId, const double* __restrict__ weight, int nidLevel,
                const uint32_t* __restrict__ feat, int n, int* __restrict__ outWord, double* __restrict__ outWeight,
                int* __restrict__ outNode) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    uint32_t d[8];
    const uint4* pd = reinterpret_cast<const uint4*>(feat + 8 * (size_t)i);
    *reinterpret_cast<uint4*>(d) = pd[0];
    *reinterpret_cast<uint4*>(d + 4) = pd[1];
    int id = 0, level = 0;
    int nid = nidLevel <= 0 ? 0 : -1;
    for (;;) {
        const int cb = childStart[id], ce = childStart[id + 1];
        if (cb == ce) break;  // leaf
        ++level;
        uint32_t key = 0xFFFFFFFFu;
        for (int c = cb + lane; c < ce; c += 32) {
            const int child = children[c];
            const uint4* cd = reinterpret_cast<const uint4*>(vdesc + 8 * (size_t)child);
            const uint32_t k = ((uint32_t)hamming8w(d, cd[0], cd[1]) << 20) | (uint32_t)(c - cb);
            key = min(key, k);
        }
        key = __reduce_min_sync(0xffffffffu, key);
        id = children[cb + (int)(key & 0xFFFFFu)];
        if (level == nidLevel) nid = id;
    }
    if (lane == 0) {
        outWord[i] = wordId[id];
        outWeight[i] = weight[id];
        outNode[i] = nid;
    }
}

struct BowSideDev {
    int n, nNodes;
    const uint32_t* desc;
    const float* angle;
    const uint8_t* valid;   // may be null (all valid)
    const int *node, *start, *feat;
};

// Top-2 of packed keys across the warp (keys are unique: they carry the list position).
__device__ __forceinline__ void warp_top2(uint32_t& k1, uint32_t& k2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        const uint32_t lo = min(k1, o1), hi = max(k1, o1);
        k2 = min(hi, min(k2, o2));
        k1 = lo;
    }
}

// ORBmatcher::SearchByBoW, node-wise join (:275-312 / :921-940) + best / second best (:313-350 / :941-971) +
// acceptance (:351-407 / :973-978).  taken[] = vpMapPointMatches[idx] != NULL / vbMatched2[idx].
__global__ void __launch_bounds__(128)
k_bow_match(BowSideDev A, BowSideDev B, int thLow, int strict, float nnratio, int nLeftB, uint8_t* taken,
            int* __restrict__ matchA, int* __restrict__ matchAR) {
    const int ia = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ia >= A.nNodes) return;
    const int node = A.node[ia];
    int lo = 0, hi = B.nNodes;   // lower_bound of `node` in B's (ascending) node list
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (B.node[mid] < node) lo = mid + 1; else hi = mid;
    }
    if (lo >= B.nNodes || B.node[lo] != node) return;
    const int bb = B.start[lo], be = B.start[lo + 1];
    for (int pa = A.start[ia]; pa < A.start[ia + 1]; pa++) {
        const int iA = A.feat[pa];
        if (A.valid && !A.valid[iA]) continue;
        uint32_t d[8];
        const uint4* pd = reinterpret_cast<const uint4*>(A.desc + 8 * (size_t)iA);
        *reinterpret_cast<uint4*>(d) = pd[0];
        *reinterpret_cast<uint4*>(d + 4) = pd[1];
        uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu, r1 = 0xFFFFFFFFu;
        for (int pb = bb + lane; pb < be; pb += 32) {
            const int iB = B.feat[pb];
            if (reinterpret_cast<volatile uint8_t*>(taken)[iB] || (B.valid && !B.valid[iB])) continue;
            const uint4* bd = reinterpret_cast<const uint4*>(B.desc + 8 * (size_t)iB);
            const uint32_t key = ((uint32_t)hamming8w(d, bd[0], bd[1]) << 20) | (uint32_t)(pb - bb);
            if (nLeftB == -1 || iB < nLeftB) {
                if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
            } else {
                r1 = min(r1, key);
            }
        }
        warp_top2(k1, k2);
        r1 = __reduce_min_sync(0xffffffffu, r1);
        const int d1 = k1 == 0xFFFFFFFFu ? 256 : (int)(k1 >> 20);
        const int d2 = k2 == 0xFFFFFFFFu ? 256 : (int)(k2 >> 20);
        const bool pass = d1 < 256 && (strict ? d1 < thLow : d1 <= thLow);
        if (pass) {
            if ((float)d1 < nnratio * (float)d2) {
                const int iB = B.feat[bb + (int)(k1 & 0xFFFFFu)];
                if (lane == 0) { matchA[iA] = iB; taken[iB] = 1; }
            }
            if (nLeftB != -1 && r1 != 0xFFFFFFFFu && (int)(r1 >> 20) <= thLow && (int)(r1 >> 20) < 256) {
                const int iB = B.feat[bb + (int)(r1 & 0xFFFFFu)];
                if (lane == 0) { matchAR[iA] = iB; taken[iB] = 1; }
            }
        }
        __syncwarp();   // lane 0's taken[] writes are visible to the whole warp for the next keyframe feature
    }
}

// rotation histogram votes (:355-372 / :983-993); bin of the left match in binOf[2*iA], right in [2*iA+1]
__global__ void k_bow_votes(int nA, const float* __restrict__ angleA, const float* __restrict__ angleB, int useHist,
                            const int* __restrict__ matchA, const int* __restrict__ matchAR, int* __restrict__ binOf,
                            int* __restrict__ hist, int* __restrict__ nmatches) {
    const int iA = blockIdx.x * blockDim.x + threadIdx.x;
    if (iA >= nA) return;
    for (int side = 0; side < 2; side++) {
        const int iB = side ? (matchAR ? matchAR[iA] : -1) : matchA[iA];
        if (iB < 0) continue;
        atomicAdd(nmatches, 1);
        if (useHist) {
            float rot = angleA[iA] - angleB[iB];
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * (1.0f / HISTO));
            if (bin == HISTO) bin = 0;
            bin = min(max(bin, 0), HISTO - 1);
            binOf[2 * iA + side] = bin;
            atomicAdd(&hist[bin], 1);
        }
    }
}

__global__ void k_bow_cull(int nA, const int* __restrict__ binOf, const int* __restrict__ hist, int* __restrict__ matchA,
                           int* __restrict__ matchAR, int* __restrict__ nmatches) {
    __shared__ int keep[3];
    if (threadIdx.x == 0) {  // ComputeThreeMaxima, ORBmatcher.cc:2336-2378
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    const int iA = blockIdx.x * blockDim.x + threadIdx.x;
    if (iA >= nA) return;
    for (int side = 0; side < 2; side++) {
        int* m = side ? matchAR : matchA;
        if (!m || m[iA] < 0) continue;
        const int b = binOf[2 * iA + side];
        if (b != keep[0] && b != keep[1] && b != keep[2]) { m[iA] = -1; atomicSub(nmatches, 1); }
    }
}

// ORBmatcher::SearchForTriangulation (:1046-1324), pinhole keyframes: no state is carried between keyframe-1 features
// (vbMatched2 is never set inside the loop of this version), so one warp serves one keyframe-1 feature: lanes over the
// keyframe-2 features of the same vocabulary node, Hamming + epipole-distance (:1162-1170) + epipolar-line
// (Pinhole::epipolarConstrain, Pinhole.cpp:196-215) gates per candidate.  `dist > bestDist` (not >=) lets a LATER
// candidate of equal distance replace the current best: key = distance << 20 | (0xFFFFF - position).
struct TriSideDev {
    int n, nNodes;
    const OrbfeKeyPoint* keys;
    const uint32_t* desc;
    const float* uright;
    const uint8_t* hasMp;
    const int *node, *start, *feat;
};
struct TriPrm {
    float F[9], ep[2];
    const float *sf2, *sigma2;
    int nLevels, onlyStereo, coarse, thLow;
    // two-camera keyframes (mpCamera2 != NULL, :1071-1095, :1160-1241): rigs = Kb8Rig[4] {ll, lr, rl, rr} on the device
    const Kb8Rig* rigs;
    const float* sigma2A;
    int nLeftA, nLeftB;
};

__global__ void __launch_bounds__(128)
k_tri_match(TriSideDev A, TriSideDev B, TriPrm P, int nfa, int* __restrict__ matches12) {
    const int pa = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pa >= nfa) return;
    int lo = 0, hi = A.nNodes;   // node list index ia with start[ia] <= pa < start[ia+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (A.start[mid] <= pa) lo = mid; else hi = mid;
    }
    const int node = A.node[lo];
    int l2 = 0, h2 = B.nNodes;
    while (l2 < h2) {
        const int mid = (l2 + h2) >> 1;
        if (B.node[mid] < node) l2 = mid + 1; else h2 = mid;
    }
    if (l2 >= B.nNodes || B.node[l2] != node) return;
    const int idx1 = A.feat[pa];
    if (A.hasMp[idx1]) return;
    const bool stereo1 = !P.rigs && A.uright && A.uright[idx1] >= 0;   // :1121 bStereo1 = !mpCamera2 && mvuRight >= 0
    if (P.onlyStereo && !stereo1) return;
    const int right1 = P.rigs && idx1 >= P.nLeftA ? 1 : 0;
    const OrbfeKeyPoint kp1 = A.keys[idx1];
    uint32_t d[8];
    const uint4* pd = reinterpret_cast<const uint4*>(A.desc + 8 * (size_t)idx1);
    *reinterpret_cast<uint4*>(d) = pd[0];
    *reinterpret_cast<uint4*>(d + 4) = pd[1];
    // epipolar line of kp1 in image 2 (Pinhole.cpp:200-202)
    const float a = kp1.x * P.F[0] + kp1.y * P.F[3] + P.F[6];
    const float b = kp1.x * P.F[1] + kp1.y * P.F[4] + P.F[7];
    const float c = kp1.x * P.F[2] + kp1.y * P.F[5] + P.F[8];
    const float den = a * a + b * b;
    const int bb = B.start[l2], be = B.start[l2 + 1];
    uint32_t key = 0xFFFFFFFFu;
    for (int pb = bb + lane; pb < be; pb += 32) {
        const int idx2 = B.feat[pb];
        if (B.hasMp[idx2]) continue;
        const bool stereo2 = !P.rigs && B.uright && B.uright[idx2] >= 0;
        if (P.onlyStereo && !stereo2) continue;
        const uint4* bd = reinterpret_cast<const uint4*>(B.desc + 8 * (size_t)idx2);
        const int dist = hamming8w(d, bd[0], bd[1]);
        if (dist > P.thLow) continue;
        const OrbfeKeyPoint kp2 = B.keys[idx2];
        if (kp2.octave < 0 || kp2.octave >= P.nLevels) continue;
        if (!stereo1 && !stereo2 && !P.rigs) {                            // :1196 ... && !pKF1->mpCamera2
            const float ex = P.ep[0] - kp2.x, ey = P.ep[1] - kp2.y;
            if (ex * ex + ey * ey < 100.0f * P.sf2[kp2.octave]) continue;
        }
        if (!P.coarse && P.rigs) {
            // KannalaBrandt8::epipolarConstrain (KannalaBrandt8.cpp:322-328) with the (bRight1, bRight2) cameras and pose
            if (kp1.octave < 0 || kp1.octave >= P.nLevels) continue;
            const int right2 = idx2 >= P.nLeftB ? 1 : 0;
            const float a1[2] = {kp1.x, kp1.y}, a2[2] = {kp2.x, kp2.y};
            float x3D[3];
            bool okTri;
            const float z = kb8_triangulate_one(P.rigs[2 * right1 + right2], a1, a2, P.sigma2A[kp1.octave], P.sigma2[kp2.octave], x3D, okTri);
            if (!(z > 0.0001f)) continue;
        } else if (!P.coarse) {
            if (den == 0) continue;
            const float num = a * kp2.x + b * kp2.y + c;
            const float dsqr = num * num / den;
            if (!((double)dsqr < 3.84 * (double)P.sigma2[kp2.octave])) continue;
        }
        key = min(key, ((uint32_t)dist << 20) | (uint32_t)(0xFFFFF - (pb - bb)));
    }
    key = __reduce_min_sync(0xffffffffu, key);
    if (lane == 0 && key != 0xFFFFFFFFu) matches12[idx1] = B.feat[bb + (0xFFFFF - (int)(key & 0xFFFFFu))];
}

__global__ void k_tri_votes(int nA, const OrbfeKeyPoint* __restrict__ keysA, const OrbfeKeyPoint* __restrict__ keysB,
                            int useHist, const int* __restrict__ m12, int* __restrict__ binOf, int* __restrict__ hist,
                            int* __restrict__ nmatches) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nA) return;
    const int j = m12[i];
    if (j < 0) return;
    atomicAdd(nmatches, 1);
    if (useHist) {
        float rot = keysA[i].angle - keysB[j].angle;
        if (rot < 0.0f) rot += 360.0f;
        int bin = (int)roundf(rot * (1.0f / HISTO));
        if (bin == HISTO) bin = 0;
        bin = min(max(bin, 0), HISTO - 1);
        binOf[2 * i] = bin;
        atomicAdd(&hist[bin], 1);
    }
}

int check_device(int device) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return bfail(ORBFE_ERR_CUDA, "no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return bfail(ORBFE_ERR_INVALID, "bad device ordinal");
    return ORBFE_OK;
}

}  // namespace

extern "C" int orbfe_vocabulary_create(int k, int L, int n_nodes, const int32_t* parent, const uint8_t* desc,
                                       const double* weight, int device, OrbfeVocabulary** out) {
    if (!out) return bfail(ORBFE_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = check_device(device);
    if (rc != ORBFE_OK) return rc;
    if (!parent || !desc || !weight || n_nodes < 2 || k < 1 || L < 1) return bfail(ORBFE_ERR_INVALID, "bad vocabulary arrays");
    if (n_nodes >= (1 << 30)) return bfail(ORBFE_ERR_CAPACITY, "vocabulary too large");
    // children in the order the nodes were added (loadFromTextFile :1390 / HKmeansStep), i.e. ascending node id
    std::vector<int> cnt(n_nodes + 1, 0), children(n_nodes - 1), wordId(n_nodes, -1);
    for (int i = 1; i < n_nodes; i++) {
        if (parent[i] < 0 || parent[i] >= n_nodes || parent[i] == i) return bfail(ORBFE_ERR_INVALID, "bad parent id");
        cnt[parent[i] + 1]++;
    }
    int maxc = 0;
    for (int i = 0; i < n_nodes; i++) { maxc = std::max(maxc, cnt[i + 1]); cnt[i + 1] += cnt[i]; }
    if (maxc >= (1 << 20)) return bfail(ORBFE_ERR_CAPACITY, "too many children per node");
    if (cnt[1] == 0) return bfail(ORBFE_ERR_INVALID, "the root has no children");
    {
        std::vector<int> fill(cnt.begin(), cnt.end() - 1);
        for (int i = 1; i < n_nodes; i++) children[fill[parent[i]]++] = i;
    }
    int w = 0;
    for (int i = 1; i < n_nodes; i++)
        if (cnt[i] == cnt[i + 1]) wordId[i] = w++;  // leaves numbered in node order (:1409-1416, createWords)
    OrbfeVocabulary* V = new OrbfeVocabulary();
    V->device = device; V->k = k; V->L = L; V->nNodes = n_nodes; V->maxChildren = maxc;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&V->desc, 32 * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc(&V->childStart, 4 * (size_t)(n_nodes + 1));
    if (e == cudaSuccess) e = cudaMalloc(&V->children, 4 * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc(&V->wordId, 4 * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc(&V->weight, 8 * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMemcpy(V->desc, desc, 32 * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(V->childStart, cnt.data(), 4 * (size_t)(n_nodes + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(V->children, children.data(), 4 * (size_t)(n_nodes - 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(V->wordId, wordId.data(), 4 * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(V->weight, weight, 8 * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        orbfe_vocabulary_destroy(V);
        return bfail(ORBFE_ERR_CUDA, "vocabulary upload", e);
    }
    *out = V;
    return ORBFE_OK;
}

extern "C" void orbfe_vocabulary_destroy(OrbfeVocabulary* V) {
    if (!V) return;
    cudaSetDevice(V->device);
    cudaFree(V->desc); cudaFree(V->childStart); cudaFree(V->children); cudaFree(V->wordId); cudaFree(V->weight);
    delete V;
}

extern "C" int orbfe_bow_transform_device(OrbfeVocabulary* V, const uint8_t* d_desc, int n, int levelsup,
                                          int32_t* d_word_id, double* d_weight, int32_t* d_node_id, void* stream) {
    if (!V || !d_word_id || !d_weight || !d_node_id || n < 0 || (n > 0 && !d_desc)) return bfail(ORBFE_ERR_INVALID, "null argument");
    if (n == 0) return ORBFE_OK;
    BCK(cudaSetDevice(V->device));
    k_bow_transform<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(V->desc, V->childStart, V->children, V->wordId, V->weight,
                                                                    V->L - levelsup, reinterpret_cast<const uint32_t*>(d_desc),
                                                                    n, d_word_id, d_weight, d_node_id);
    BCK(cudaGetLastError());
    return ORBFE_OK;
}

// ---------------------------------------------------------------------------------------------
// BowVector / FeatureVector folding on the device: transform(features, v, fv, levelsup) (TemplatedVocabulary.h:1125-1197)
// after the per-feature descent, for the TF_IDF / L1_NORM vocabulary ORB-SLAM3 ships (ORBvoc.txt).  DBoW2 folds into two
// std::maps in feature order: v[word] += weight (BowVector::addWeight), fv[node].push_back(i), then v.normalize(L1)
// sums |value| in key order and divides.  Here, one CTA per frame: the (word, i) and (node, i) keys are sorted in shared
// memory (bitonic; a key is unique, so the order inside a word / node is the feature order), one thread per word adds its
// weights in that order (the same sequence of double additions as the map's +=), ONE thread adds the norm in key order
// (the only way to reproduce the reference's rounding), every thread divides.
constexpr int FOLD_CAP = 4096;     // features per frame (shared-memory sort)
constexpr int FOLD_THREADS = 256;

__device__ __forceinline__ void fold_sort(unsigned long long* key, int n2) {
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += FOLD_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = key[i], b = key[l];
                    if ((a > b) == ((i & k) == 0)) { key[i] = b; key[l] = a; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(FOLD_THREADS)
k_bow_fold(const int32_t* __restrict__ word, const double* __restrict__ weight, const int32_t* __restrict__ node,
           const int32_t* __restrict__ fstart, int cap, uint32_t* __restrict__ bowWord, double* __restrict__ bowVal,
           int32_t* __restrict__ nBow, uint32_t* __restrict__ fvNode, int32_t* __restrict__ fvStart, int32_t* __restrict__ fvFeat,
           int32_t* __restrict__ nFv) {
    __shared__ unsigned long long key[FOLD_CAP];
    __shared__ unsigned short headPos[FOLD_CAP];   // slot -> position of its first key
    __shared__ int s_n, s_slots;
    __shared__ double s_norm;
    const int b = blockIdx.x, f0 = fstart[b], n = min(fstart[b + 1] - f0, min(cap, FOLD_CAP));
    const size_t o = (size_t)b * cap;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int pass = 0; pass < 2; pass++) {          // 0: BowVector (keys = word), 1: FeatureVector (keys = node)
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n2; i += FOLD_THREADS) {
            unsigned long long k = ~0ull;
            if (i < n && weight[f0 + i] > 0.0)      // stopped words are skipped in both maps (:1170)
                k = ((unsigned long long)(uint32_t)(pass ? node[f0 + i] : word[f0 + i]) << 32) | (uint32_t)i;
            key[i] = k;
            if (k != ~0ull) atomicAdd(&s_n, 1);
        }
        __syncthreads();
        fold_sort(key, n2);
        const int m = s_n;                          // valid keys are the first m
        // segment heads -> slots (block scan by one pass of warp ballots is overkill for <= 4096: two-level count)
        if (threadIdx.x == 0) s_slots = 0;
        __syncthreads();
        for (int base = 0; base < m; base += FOLD_THREADS) {
            const int i = base + threadIdx.x;
            const bool head = i < m && (i == 0 || (key[i] >> 32) != (key[i - 1] >> 32));
            // ordered slot numbers: count the heads before i inside this chunk with ballots
            const unsigned bal = __ballot_sync(0xffffffffu, head);
            __shared__ int wcnt[FOLD_THREADS / 32];
            if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = __popc(bal);
            __syncthreads();
            int before = s_slots;
            for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += wcnt[w];
            if (head) headPos[before + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = (unsigned short)i;
            __syncthreads();
            if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < FOLD_THREADS / 32; w++) t += wcnt[w]; s_slots += t; }
            __syncthreads();
        }
        const int slots = s_slots;
        if (pass == 0) {
            for (int sl = threadIdx.x; sl < slots; sl += FOLD_THREADS) {
                const int p0 = headPos[sl], p1 = sl + 1 < slots ? headPos[sl + 1] : m;
                double v = weight[f0 + (int)(uint32_t)key[p0]];                    // insert(value_type(id, v))
                for (int p = p0 + 1; p < p1; p++) v += weight[f0 + (int)(uint32_t)key[p]];   // vit->second += v, feature order
                bowWord[o + sl] = (uint32_t)(key[p0] >> 32);
                bowVal[o + sl] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) {                 // BowVector::normalize(L1), BowVector.cpp:62-84: key order, one accumulator
                double norm = 0.0;
                for (int sl = 0; sl < slots; sl++) norm += fabs(bowVal[o + sl]);
                s_norm = norm;
                nBow[b] = slots;
            }
            __syncthreads();
            const double norm = s_norm;
            if (norm > 0.0)
                for (int sl = threadIdx.x; sl < slots; sl += FOLD_THREADS) bowVal[o + sl] = bowVal[o + sl] / norm;
        } else {
            for (int sl = threadIdx.x; sl < slots; sl += FOLD_THREADS) {
                fvNode[o + sl] = (uint32_t)(key[headPos[sl]] >> 32);
                fvStart[(size_t)b * (cap + 1) + sl] = headPos[sl];
            }
            for (int i = threadIdx.x; i < m; i += FOLD_THREADS) fvFeat[o + i] = (int32_t)(uint32_t)key[i];
            if (threadIdx.x == 0) { fvStart[(size_t)b * (cap + 1) + slots] = m; nFv[b] = slots; }
        }
        __syncthreads();
    }
}

extern "C" int orbfe_bow_fold_device(const int32_t* d_word_id, const double* d_weight, const int32_t* d_node_id,
                                     const int32_t* d_frame_start, int B, int capacity, uint32_t* d_bow_word,
                                     double* d_bow_value, int32_t* d_n_bow, uint32_t* d_fv_node, int32_t* d_fv_start,
                                     int32_t* d_fv_feat, int32_t* d_n_fv, void* stream) {
    if (B <= 0) return ORBFE_OK;
    if (!d_word_id || !d_weight || !d_node_id || !d_frame_start || !d_bow_word || !d_bow_value || !d_n_bow || !d_fv_node ||
        !d_fv_start || !d_fv_feat || !d_n_fv)
        return bfail(ORBFE_ERR_INVALID, "null argument");
    if (capacity <= 0 || capacity > FOLD_CAP) return bfail(ORBFE_ERR_CAPACITY, "bow fold: at most 4096 features per frame");
    k_bow_fold<<<B, FOLD_THREADS, 0, (cudaStream_t)stream>>>(d_word_id, d_weight, d_node_id, d_frame_start, capacity, d_bow_word,
                                                          d_bow_value, d_n_bow, d_fv_node, d_fv_start, d_fv_feat, d_n_fv);
    BCK(cudaGetLastError());
    return ORBFE_OK;
}

extern "C" int orbfe_bow_transform(OrbfeVocabulary* V, const uint8_t* desc, int n, int levelsup, int32_t* word_id,
                                   double* weight, int32_t* node_id) {
    if (!V || !word_id || !weight || !node_id || n < 0 || (n > 0 && !desc)) return bfail(ORBFE_ERR_INVALID, "null argument");
    if (n == 0) return ORBFE_OK;
    OrbfeStage S;
    const size_t iD = S.in(desc, 32 * (size_t)n);
    const size_t oWt = S.out(weight, 8 * (size_t)n), oW = S.out(word_id, 4 * (size_t)n), oN = S.out(node_id, 4 * (size_t)n);
    BCK(S.commit(V->device));
    BCK(S.upload());
    int rc = orbfe_bow_transform_device(V, S.ptr<uint8_t>(iD), n, levelsup, S.ptr<int32_t>(oW), S.ptr<double>(oWt),
                                        S.ptr<int32_t>(oN), S.stream());
    if (rc != ORBFE_OK) return rc;
    BCK(S.download());
    return ORBFE_OK;
}

extern "C" int orbfe_search_by_bow(const OrbfeBowSide* a, const OrbfeBowSide* b, int th_low, int strict, float nnratio,
                                   int check_orientation, int n_left_b, int32_t* match_a, int32_t* match_a_right,
                                   int device) {
    int rc = check_device(device);
    if (rc != ORBFE_OK) return rc;
    if (!a || !b || !match_a) return bfail(ORBFE_ERR_INVALID, "null argument");
    if (a->n < 0 || b->n < 0 || a->fv.n_nodes < 0 || b->fv.n_nodes < 0) return bfail(ORBFE_ERR_INVALID, "bad sizes");
    if (n_left_b != -1 && !match_a_right) return bfail(ORBFE_ERR_INVALID, "fisheye frames need match_a_right");
    for (int i = 0; i < a->n; i++) { match_a[i] = -1; if (match_a_right) match_a_right[i] = -1; }
    if (a->n == 0 || b->n == 0 || a->fv.n_nodes == 0 || b->fv.n_nodes == 0) return 0;
    if (!a->desc || !b->desc || !a->fv.node_id || !a->fv.start || !a->fv.feat || !b->fv.node_id || !b->fv.start ||
        !b->fv.feat || (check_orientation && (!a->angle || !b->angle)))
        return bfail(ORBFE_ERR_INVALID, "missing array");
    const int nfa = a->fv.start[a->fv.n_nodes], nfb = b->fv.start[b->fv.n_nodes];
    if (nfa < 0 || nfb < 0 || nfb >= (1 << 20)) return bfail(ORBFE_ERR_CAPACITY, "feature vector too long");
    const bool right = n_left_b != -1;
    OrbfeStage S;
    struct Lay { size_t desc, angle, valid, node, start, feat; } la, lb;
    auto lay = [&](const OrbfeBowSide* s, int nf, Lay& l) {
        l.desc = S.in(s->desc, 32 * (size_t)s->n);
        l.angle = S.in(s->angle, s->angle ? 4 * (size_t)s->n : 0);
        l.valid = S.in(s->valid, s->valid ? (size_t)s->n : 0);
        l.node = S.in(s->fv.node_id, 4 * (size_t)s->fv.n_nodes);
        l.start = S.in(s->fv.start, 4 * (size_t)(s->fv.n_nodes + 1));
        l.feat = S.in(s->fv.feat, 4 * (size_t)nf);
    };
    lay(a, nfa, la);
    lay(b, nfb, lb);
    const size_t wTaken = S.work((size_t)b->n), wBin = S.work(8 * (size_t)a->n), wHist = S.work(4 * (HISTO + 2));
    int nmatches = 0;
    const size_t oM = S.out(match_a, 4 * (size_t)a->n), oR = S.out(match_a_right, right ? 4 * (size_t)a->n : 0);
    const size_t oN = S.out(&nmatches, 4);
    BCK(S.commit(device));
    cudaStream_t st = S.stream();
    BCK(S.upload());
    auto bind = [&](const OrbfeBowSide* s, const Lay& l) {
        BowSideDev d;
        d.n = s->n; d.nNodes = s->fv.n_nodes; d.desc = S.ptr<uint32_t>(l.desc);
        d.angle = s->angle ? S.ptr<float>(l.angle) : nullptr; d.valid = s->valid ? S.ptr<uint8_t>(l.valid) : nullptr;
        d.node = S.ptr<int>(l.node); d.start = S.ptr<int>(l.start); d.feat = S.ptr<int>(l.feat);
        return d;
    };
    const BowSideDev A = bind(a, la), B = bind(b, lb);
    int* dM = S.ptr<int>(oM);
    int* dR = right ? S.ptr<int>(oR) : nullptr;
    BCK(cudaMemsetAsync(S.ptr<uint8_t>(wTaken), 0, (size_t)b->n, st));
    BCK(cudaMemsetAsync(dM, 0xFF, 4 * (size_t)a->n, st));
    if (right) BCK(cudaMemsetAsync(dR, 0xFF, 4 * (size_t)a->n, st));
    BCK(cudaMemsetAsync(S.ptr<int>(wHist), 0, 4 * (HISTO + 2), st));
    BCK(cudaMemsetAsync(S.ptr<int>(oN), 0, 4, st));
    k_bow_match<<<(A.nNodes + 3) / 4, 128, 0, st>>>(A, B, th_low, strict ? 1 : 0, nnratio, n_left_b, S.ptr<uint8_t>(wTaken), dM, dR);
    const int g = (a->n + 255) / 256;
    k_bow_votes<<<g, 256, 0, st>>>(a->n, A.angle, B.angle, check_orientation ? 1 : 0, dM, dR, S.ptr<int>(wBin), S.ptr<int>(wHist),
                                   S.ptr<int>(oN));
    if (check_orientation) k_bow_cull<<<g, 256, 0, st>>>(a->n, S.ptr<int>(wBin), S.ptr<int>(wHist), dM, dR, S.ptr<int>(oN));
    BCK(cudaGetLastError());
    BCK(S.download());
    return nmatches;
}

extern "C" int orbfe_search_for_triangulation(const OrbfeTriSide* kf1, const OrbfeTriSide* kf2, const OrbfeTriParams* prm,
                                              int32_t* matches12, int device) {
    int rc = check_device(device);
    if (rc != ORBFE_OK) return rc;
    if (!kf1 || !kf2 || !prm || !matches12) return bfail(ORBFE_ERR_INVALID, "null argument");
    if (kf1->n < 0 || kf2->n < 0 || kf1->fv.n_nodes < 0 || kf2->fv.n_nodes < 0 || prm->n_levels <= 0)
        return bfail(ORBFE_ERR_INVALID, "bad sizes");
    for (int i = 0; i < kf1->n; i++) matches12[i] = -1;
    if (kf1->n == 0 || kf2->n == 0 || kf1->fv.n_nodes == 0 || kf2->fv.n_nodes == 0) return 0;
    if (!kf1->keys || !kf1->desc || !kf1->has_map_point || !kf2->keys || !kf2->desc || !kf2->has_map_point ||
        !kf1->fv.node_id || !kf1->fv.start || !kf1->fv.feat || !kf2->fv.node_id || !kf2->fv.start || !kf2->fv.feat ||
        !prm->scale_factors2 || !prm->level_sigma2_2)
        return bfail(ORBFE_ERR_INVALID, "missing array");
    const int nfa = kf1->fv.start[kf1->fv.n_nodes], nfb = kf2->fv.start[kf2->fv.n_nodes];
    if (nfa < 0 || nfb < 0 || nfb >= (1 << 20)) return bfail(ORBFE_ERR_CAPACITY, "feature vector too long");
    OrbfeStage S;
    struct Lay { size_t keys, desc, ur, mp, node, start, feat; } la, lb;
    auto lay = [&](const OrbfeTriSide* s, int nf, Lay& l) {
        l.keys = S.in(s->keys, sizeof(OrbfeKeyPoint) * (size_t)s->n);
        l.desc = S.in(s->desc, 32 * (size_t)s->n);
        l.ur = S.in(s->uright, s->uright ? 4 * (size_t)s->n : 0);
        l.mp = S.in(s->has_map_point, (size_t)s->n);
        l.node = S.in(s->fv.node_id, 4 * (size_t)s->fv.n_nodes);
        l.start = S.in(s->fv.start, 4 * (size_t)(s->fv.n_nodes + 1));
        l.feat = S.in(s->fv.feat, 4 * (size_t)nf);
    };
    lay(kf1, nfa, la);
    lay(kf2, nfb, lb);
    const size_t iSf = S.in(prm->scale_factors2, 4 * (size_t)prm->n_levels), iS2 = S.in(prm->level_sigma2_2, 4 * (size_t)prm->n_levels);
    Kb8Rig rigs[4];
    size_t iRig = 0, iS1 = 0;
    if (prm->rig) {
        if (!prm->rig->level_sigma2_1) return bfail(ORBFE_ERR_INVALID, "two-camera search: level_sigma2_1 missing");
        for (int k = 0; k < 4; k++) {
            const OrbfeTriCameraPair& c = prm->rig->pair[k];
            for (int i = 0; i < 8; i++) { rigs[k].c1.p[i] = c.params1[i]; rigs[k].c2.p[i] = c.params2[i]; }
            rigs[k].c1.precision = c.precision1; rigs[k].c2.precision = c.precision2;
            for (int i = 0; i < 9; i++) rigs[k].R12[i] = c.R12[i];
            for (int i = 0; i < 3; i++) rigs[k].t12[i] = c.t12[i];
        }
        iRig = S.in(rigs, sizeof rigs);
        iS1 = S.in(prm->rig->level_sigma2_1, 4 * (size_t)prm->n_levels);
    }
    const size_t wBin = S.work(8 * (size_t)kf1->n), wHist = S.work(4 * (HISTO + 2));
    int nmatches = 0;
    const size_t oM = S.out(matches12, 4 * (size_t)kf1->n), oN = S.out(&nmatches, 4);
    BCK(S.commit(device));
    cudaStream_t st = S.stream();
    BCK(S.upload());
    auto bind = [&](const OrbfeTriSide* s, const Lay& l) {
        TriSideDev d;
        d.n = s->n; d.nNodes = s->fv.n_nodes; d.keys = S.ptr<OrbfeKeyPoint>(l.keys); d.desc = S.ptr<uint32_t>(l.desc);
        d.uright = s->uright ? S.ptr<float>(l.ur) : nullptr; d.hasMp = S.ptr<uint8_t>(l.mp);
        d.node = S.ptr<int>(l.node); d.start = S.ptr<int>(l.start); d.feat = S.ptr<int>(l.feat);
        return d;
    };
    const TriSideDev A = bind(kf1, la), B = bind(kf2, lb);
    TriPrm P;
    for (int i = 0; i < 9; i++) P.F[i] = prm->f12[i];
    P.ep[0] = prm->epipole[0]; P.ep[1] = prm->epipole[1];
    P.sf2 = S.ptr<float>(iSf); P.sigma2 = S.ptr<float>(iS2); P.nLevels = prm->n_levels;
    P.onlyStereo = prm->only_stereo; P.coarse = prm->coarse; P.thLow = prm->th_low;
    P.rigs = prm->rig ? S.ptr<Kb8Rig>(iRig) : nullptr;
    P.sigma2A = prm->rig ? S.ptr<float>(iS1) : nullptr;
    P.nLeftA = prm->rig ? prm->rig->n_left1 : -1;
    P.nLeftB = prm->rig ? prm->rig->n_left2 : -1;
    int* dM = S.ptr<int>(oM);
    BCK(cudaMemsetAsync(dM, 0xFF, 4 * (size_t)kf1->n, st));
    BCK(cudaMemsetAsync(S.ptr<int>(wHist), 0, 4 * (HISTO + 2), st));
    BCK(cudaMemsetAsync(S.ptr<int>(oN), 0, 4, st));
    k_tri_match<<<(nfa + 3) / 4, 128, 0, st>>>(A, B, P, nfa, dM);
    const int g = (kf1->n + 255) / 256;
    k_tri_votes<<<g, 256, 0, st>>>(kf1->n, A.keys, B.keys, prm->check_orientation ? 1 : 0, dM, S.ptr<int>(wBin), S.ptr<int>(wHist),
                                   S.ptr<int>(oN));
    if (prm->check_orientation) k_bow_cull<<<g, 256, 0, st>>>(kf1->n, S.ptr<int>(wBin), S.ptr<int>(wHist), dM, nullptr, S.ptr<int>(oN));
    BCK(cudaGetLastError());
    BCK(S.download());
    return nmatches;
}
