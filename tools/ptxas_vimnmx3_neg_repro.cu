// Reproducer of the sm_100a miscompile found in round 1 (nvcc 12.9.86): on a B200 the plain-loop FAST score
// `ref_best` and the `max(best, max(mn, -mx))` formulation (fc_arc_best of the first version, kept below as V1)
// return wrong values on the DEVICE while V2 (negate once, outside the min/max network) and V4 (packed s16x2)
// are correct; on the host all four agree.  Output on B200: "V1 bad=7785 V2 bad=0 ref-on-device bad=7769 V4 bad=0".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../orb-slam3_byzyh_b200/csrc/fast_core.h"
constexpr int P = 72, ROWS = 22, NT = 8192;

// V1: the first formulation (doubling scheme, running max over max(mn9, -mx9)) -- wrong on the device.
template <int PITCH>
__host__ __device__ inline int fc_arc_best(const uint8_t* p) {
    int d[16];
    const int v = p[0];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - (int)p[FC_RING_DX(k) + FC_RING_DY(k) * PITCH];
    int mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = max(best, max(mn9, -mx9));
    }
    return best;
}

__host__ __device__ inline int ref_best(const uint8_t* p) {   // plain loops
    const int dx[16] = {0,1,2,3,3,3,2,1,0,-1,-2,-3,-3,-3,-2,-1};
    const int dy[16] = {3,3,2,1,0,-1,-2,-3,-3,-3,-2,-1,0,1,2,3};
    int d[16];
    for (int k = 0; k < 16; k++) d[k] = (int)p[0] - (int)p[dx[k] + dy[k] * P];
    int best = 0;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) { int v = d[(s + k) & 15]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best;
}
// V2: A = max_k min9_k, B = min_k max9_k, best = max(0, A, -B)
template <int PITCH> __device__ __forceinline__ int v2(const uint8_t* p) {
    int d[16];
    const int v = p[0];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - (int)p[FC_RING_DX(k) + FC_RING_DY(k) * PITCH];
    int mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn3[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
        mx3[k] = max(max(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
    }
    int A = -1000, B = 1000;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        A = max(A, min(min(mn3[k], mn3[(k + 3) & 15]), mn3[(k + 6) & 15]));
        B = min(B, max(max(mx3[k], mx3[(k + 3) & 15]), mx3[(k + 6) & 15]));
    }
    return max(0, max(A, -B));
}
// V4: two horizontally adjacent pixels packed as s16x2 (lo = pixel x, hi = pixel x+1)
template <int PITCH> __device__ __forceinline__ unsigned v4(const uint8_t* p) {
    unsigned e[16];
    const unsigned c = (unsigned)p[0] | ((unsigned)p[1] << 16);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint8_t* q = p + FC_RING_DX(k) + FC_RING_DY(k) * PITCH;
        e[k] = __vsub2(c, (unsigned)q[0] | ((unsigned)q[1] << 16));
    }
    unsigned mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn3[k] = __vimin3_s16x2(e[k], e[(k + 1) & 15], e[(k + 2) & 15]);
        mx3[k] = __vimax3_s16x2(e[k], e[(k + 1) & 15], e[(k + 2) & 15]);
    }
    unsigned A = 0x80008000u, B = 0x7fff7fffu;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        A = __vmaxs2(A, __vimin3_s16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]));
        B = __vmins2(B, __vimax3_s16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]));
    }
    return __vimax3_s16x2(A, __vneg2(B), 0u);
}
__global__ void k(const uint8_t* tiles, int* o1, int* o2, int* o3, unsigned* o4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NT) return;
    const uint8_t* p = tiles + (size_t)i * P * ROWS + 10 * P + 30;
    o1[i] = fc_arc_best<P>(p);
    o2[i] = v2<P>(p);
    o3[i] = ref_best(p);
    o4[i] = v4<P>(p);
}
int main() {
    std::vector<uint8_t> h((size_t)NT * P * ROWS);
    for (size_t i = 0; i < h.size(); i++) { int m = (i / (P * ROWS)) % 4; h[i] = m == 0 ? rand() % 256 : m == 1 ? 100 + rand() % 40 : m == 2 ? (rand() % 8 ? 30 + rand() % 10 : 200 + rand() % 50) : 120 + rand() % 12; }
    uint8_t* d; int *d1, *d2, *d3; unsigned* d4;
    cudaMalloc(&d, h.size()); cudaMalloc(&d1, NT * 4); cudaMalloc(&d2, NT * 4); cudaMalloc(&d3, NT * 4); cudaMalloc(&d4, NT * 4);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    k<<<NT / 128, 128>>>(d, d1, d2, d3, d4);
    std::vector<int> o1(NT), o2(NT), o3(NT); std::vector<unsigned> o4(NT);
    cudaMemcpy(o1.data(), d1, NT * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o2.data(), d2, NT * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o3.data(), d3, NT * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o4.data(), d4, NT * 4, cudaMemcpyDeviceToHost);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    int b1 = 0, b2 = 0, b3 = 0, b4 = 0, nz = 0;
    for (int i = 0; i < NT; i++) {
        const uint8_t* p = h.data() + (size_t)i * P * ROWS + 10 * P + 30;
        int r = ref_best(p), r1 = ref_best(p + 1);
        nz += r > 7;
        b1 += o1[i] != r; b2 += o2[i] != r; b3 += o3[i] != r;
        int lo = (int)(short)(o4[i] & 0xffff), hi = (int)(short)(o4[i] >> 16);
        if (lo != r || hi != r1) { if (b4 < 4) printf("v4 mismatch %d: dev (%d,%d) host (%d,%d)\n", i, lo, hi, r, r1); b4++; }
    }
    printf("of %d (corners>7: %d): V1(current) bad=%d  V2(3-wise) bad=%d  ref-on-device bad=%d  V4(s16x2) bad=%d\n", NT, nz, b1, b2, b3, b4);
}
