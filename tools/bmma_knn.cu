// tools/bmma_knn.cu -- one-page experiment: is the 256-bit Hamming kNN-2 a job for the tensor
// cores?  mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc gives popc(a & b) for 16 x 8 descriptor pairs per
// warp instruction (k = 256 bits = one ORB descriptor); hamming(a, b) = popc(a) + popc(b) - 2 popc(a & b).
// Compares, for nq queries x nt train descriptors (one 8-GPU shard of BASELINE config 5: 2000 x 125 000),
//   (a) the scalar kernel of csrc/match.cu (5 POPC + 14 LOP3 per pair, two queries per thread), re-stated here, and
//   (b) a binary-MMA kernel: 64 queries per warp (four A fragments), train descriptors staged in shared memory, the
//       top-2 (distance << 23 | index) keys kept per accumulator row and merged over the quad at the end,
// checks that the two best-two tables are identical, and prints pairs/s.   nvcc -O3 -arch sm_100a tools/bmma_knn.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int KEY_SHIFT = 23;

__device__ __forceinline__ void top2_insert(uint32_t& b0, uint32_t& b1, uint32_t k) {
    const uint32_t hi = max(b0, k);
    b0 = min(b0, k);
    b1 = min(b1, hi);
}
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}
__device__ __forceinline__ int hamming256_csa(const uint32_t* a, const uint32_t* b) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a[i] ^ b[i];
    uint32_t s1, c1, s2, c2, s3, c3;
    csa(x[0], x[1], x[2], s1, c1);
    csa(x[3], x[4], x[5], s2, c2);
    csa(x[6], s1, s2, s3, c3);
    return (__popc(s3) + __popc(x[7])) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

// (a) scalar: thread = 2 queries, CTA = 256 queries x one chunk of train descriptors; out[q][chunk][2]
__global__ void __launch_bounds__(128) k_scalar(const uint32_t* __restrict__ query, int nq, const uint32_t* __restrict__ train,
                                                int nt, int chunk, int nchunks, uint32_t* __restrict__ partial) {
    __shared__ __align__(16) uint32_t tile[256 * 8];
    const int q0 = blockIdx.x * 256 + threadIdx.x;
    uint32_t qv[2][8], b0[2], b1[2];
    for (int r = 0; r < 2; r++) {
        const int q = min(q0 + r * 128, nq - 1);
        for (int i = 0; i < 8; i++) qv[r][i] = query[8 * (size_t)q + i];
        b0[r] = b1[r] = KEY_NONE;
    }
    const int c0 = blockIdx.y * chunk, c1 = min(c0 + chunk, nt);
    for (int t0 = c0; t0 < c1; t0 += 256) {
        const int cnt = min(256, c1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += 128) reinterpret_cast<uint4*>(tile)[i] = reinterpret_cast<const uint4*>(train + 8 * (size_t)t0)[i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            uint32_t tv[8];
            *reinterpret_cast<uint4*>(tv) = reinterpret_cast<const uint4*>(tile)[2 * j];
            *reinterpret_cast<uint4*>(tv + 4) = reinterpret_cast<const uint4*>(tile)[2 * j + 1];
#pragma unroll
            for (int r = 0; r < 2; r++) top2_insert(b0[r], b1[r], ((uint32_t)hamming256_csa(qv[r], tv) << KEY_SHIFT) | (uint32_t)(t0 + j));
        }
    }
    for (int r = 0; r < 2; r++) {
        const int q = q0 + r * 128;
        if (q < nq) { partial[((size_t)q * nchunks + blockIdx.y) * 2] = b0[r]; partial[((size_t)q * nchunks + blockIdx.y) * 2 + 1] = b1[r]; }
    }
}

// (b) binary MMA.  Warp = 64 queries (4 A fragments); CTA = 4 warps = 256 queries x one chunk of train descriptors.
constexpr int MT = 128;   // train descriptors staged per iteration
__global__ void __launch_bounds__(128) k_bmma(const uint32_t* __restrict__ query, int nq, const uint32_t* __restrict__ train, int nt,
                                              int chunk, int nchunks, uint32_t* __restrict__ partial) {
    __shared__ __align__(16) uint32_t tile[MT * 8];
    __shared__ int tpop[MT];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int qw = blockIdx.x * 256 + wid * 64;
    uint32_t a[4][4];
    int pa[4][2];
    uint32_t b0[4][2], b1[4][2];
#pragma unroll
    for (int f = 0; f < 4; f++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int q = min(qw + 16 * f + g + 8 * h, nq - 1);
            a[f][h] = query[8 * (size_t)q + t];            // a0 / a1: row g / g+8, bits 32t .. 32t+31
            a[f][2 + h] = query[8 * (size_t)q + 4 + t];    // a2 / a3: bits 128 + 32t ..
            int p = __popc(a[f][h]) + __popc(a[f][2 + h]);
            p += __shfl_xor_sync(0xffffffffu, p, 1);
            p += __shfl_xor_sync(0xffffffffu, p, 2);
            pa[f][h] = p;
            b0[f][h] = b1[f][h] = KEY_NONE;
        }
    }
    const int c0 = blockIdx.y * chunk, c1 = min(c0 + chunk, nt);
    for (int t0 = c0; t0 < c1; t0 += MT) {
        const int cnt = min(MT, c1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < MT * 2; i += 128)
            reinterpret_cast<uint4*>(tile)[i] = i < cnt * 2 ? reinterpret_cast<const uint4*>(train + 8 * (size_t)t0)[i] : make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (int i = threadIdx.x; i < MT; i += 128) {
            int p = 0;
            for (int k = 0; k < 8; k++) p += __popc(tile[8 * i + k]);
            tpop[i] = i < cnt ? p : 100000;                // padding columns can never win
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < MT; j += 8) {
            const uint32_t bb0 = tile[8 * (j + g) + t], bb1 = tile[8 * (j + g) + 4 + t];     // col g of this 8-wide slab
            const int pb0 = tpop[j + 2 * t], pb1 = tpop[j + 2 * t + 1];
            const uint32_t i0 = (uint32_t)(t0 + j + 2 * t), i1 = i0 + 1;
#pragma unroll
            for (int f = 0; f < 4; f++) {
                int c[4] = {0, 0, 0, 0};
                asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                             : "r"(a[f][0]), "r"(a[f][1]), "r"(a[f][2]), "r"(a[f][3]), "r"(bb0), "r"(bb1));
                // c0: (g, 2t)  c1: (g, 2t+1)  c2: (g+8, 2t)  c3: (g+8, 2t+1)
                top2_insert(b0[f][0], b1[f][0], ((uint32_t)(pa[f][0] + pb0 - 2 * c[0]) << KEY_SHIFT) | i0);
                top2_insert(b0[f][0], b1[f][0], ((uint32_t)(pa[f][0] + pb1 - 2 * c[1]) << KEY_SHIFT) | i1);
                top2_insert(b0[f][1], b1[f][1], ((uint32_t)(pa[f][1] + pb0 - 2 * c[2]) << KEY_SHIFT) | i0);
                top2_insert(b0[f][1], b1[f][1], ((uint32_t)(pa[f][1] + pb1 - 2 * c[3]) << KEY_SHIFT) | i1);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t x0 = b0[f][h], x1 = b1[f][h];
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {      // merge over the four threads of the quad (same row, different columns)
                const uint32_t o0 = __shfl_xor_sync(0xffffffffu, x0, o), o1 = __shfl_xor_sync(0xffffffffu, x1, o);
                top2_insert(x0, x1, o0);
                top2_insert(x0, x1, o1);
            }
            const int q = qw + 16 * f + g + 8 * h;
            if (t == 0 && q < nq) { partial[((size_t)q * nchunks + blockIdx.y) * 2] = x0 >= (100000u << KEY_SHIFT) ? KEY_NONE : x0;
                                    partial[((size_t)q * nchunks + blockIdx.y) * 2 + 1] = x1 >= (100000u << KEY_SHIFT) ? KEY_NONE : x1; }
        }
}

int main(int argc, char** argv) {
    const int nq = argc > 1 ? atoi(argv[1]) : 2000, nt = argc > 2 ? atoi(argv[2]) : 125000;
    std::vector<uint32_t> q(8 * (size_t)nq), tr(8 * (size_t)nt);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    for (auto& v : q) v = rnd();
    for (auto& v : tr) v = rnd();
    for (int i = 0; i < 8; i++) { tr[8 * 777 + i] = q[8 * 5 + i]; tr[8 * 90001 + i] = q[8 * 5 + i]; }   // ties: lower index first
    uint32_t *dq, *dt, *p1, *p2;
    const int chunk = 2048, nch = (nt + chunk - 1) / chunk;
    CK(cudaMalloc(&dq, q.size() * 4)); CK(cudaMalloc(&dt, tr.size() * 4));
    CK(cudaMalloc(&p1, (size_t)nq * nch * 8)); CK(cudaMalloc(&p2, (size_t)nq * nch * 8));
    CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice));
    dim3 grid((nq + 255) / 256, nch);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms1 = 0, ms2 = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); for (int i = 0; i < 10; i++) k_scalar<<<grid, 128>>>(dq, nq, dt, nt, chunk, nch, p1); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms1, e0, e1);
        cudaEventRecord(e0); for (int i = 0; i < 10; i++) k_bmma<<<grid, 128>>>(dq, nq, dt, nt, chunk, nch, p2); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms2, e0, e1);
    }
    CK(cudaGetLastError());
    std::vector<uint32_t> h1((size_t)nq * nch * 2), h2(h1.size());
    CK(cudaMemcpy(h1.data(), p1, h1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), p2, h2.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < h1.size(); i++) bad += h1[i] != h2[i];
    const double pairs = (double)nq * nt;
    printf("nq %d nt %d: scalar (5 POPC + 14 LOP3 per pair) %.3f ms = %.1f G pairs/s | bmma m16n8k256 and.popc %.3f ms = %.1f G pairs/s | "
           "best-two tables %s (%zu of %zu entries differ)\n", nq, nt, ms1 / 10, pairs / (ms1 / 10) / 1e6, ms2 / 10, pairs / (ms2 / 10) / 1e6,
           bad ? "DIFFER" : "identical", bad, h1.size());
    return 0;
}
