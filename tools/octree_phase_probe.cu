// tools/octree_phase_probe.cu -- where one DistributeOctTree CTA (csrc/octree_core.h) spends its cycles.
// Stand-alone probe, not part of the product: random candidates of a 752x480 level 0 (n points, N = 217 features, two
// roots), one CTA of 256 threads, clock64() accumulated per stage by thread 0.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/octree_phase_probe tools/octree_phase_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ long long g_prof[8];
__device__ long long g_last;
__device__ int g_cnt[8];
__device__ __forceinline__ void oc_mark_dev(int slot) {
    if (threadIdx.x == 0) {
        const long long t = clock64();
        g_prof[slot] += t - g_last;
        g_cnt[slot]++;
        g_last = t;
    }
}
#if defined(__CUDA_ARCH__)
#define OC_MARK(slot) oc_mark_dev(slot)
#else
#define OC_MARK(slot) ((void)0)
#endif
#include "../orb-slam3_byzyh_b200/csrc/octree_core.h"

__global__ void __launch_bounds__(256)
k_probe(const uint32_t* pk, uint32_t* pnode, int n, int width, int height, int nIni, float hX, int N, int M, int* outn) {
    extern __shared__ __align__(16) char smem[];
    OcWork w;
    oc_carve(w, smem, M);
    w.pk = pk; w.pnode = pnode; w.n = n;
    __shared__ int s_outn;
    if (threadIdx.x == 0) g_last = clock64();
    __syncthreads();
    oc_distribute(w, width, height, nIni, hX, N, w.cc, &s_outn, w.cpos);
    if (threadIdx.x == 0) *outn = s_outn;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 3000, N = argc > 2 ? atoi(argv[2]) : 217;
    const int width = 752 - 32, height = 480 - 32;   // level-0 window, roughly
    const int nIni = 2;
    const float hX = (float)width / nIni;
    const int M = (N + 3 > 4 * nIni ? N + 3 : 4 * nIni) + 1;
    std::vector<uint32_t> pk(n);
    srand(7);
    for (int i = 0; i < n; i++) pk[i] = OC_PACK(rand() % width, rand() % height, 8 + rand() % 100);
    uint32_t *d_pk, *d_pn; int* d_out;
    cudaMalloc(&d_pk, n * 4); cudaMalloc(&d_pn, n * 4); cudaMalloc(&d_out, 4);
    cudaMemcpy(d_pk, pk.data(), n * 4, cudaMemcpyHostToDevice);
    const size_t sm = oc_shared_bytes(M) + 64;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[8] = {"roots", "child populations", "phase-1 rebuild", "phase-2: copies + new list", "relabel", "winners", "phase-2: sort", "phase-2: divisions"};
    for (int rep = 0; rep < 3; rep++) {
        long long z[8] = {0}; int zc[8] = {0};
        cudaMemcpyToSymbol(g_prof, z, sizeof(z)); cudaMemcpyToSymbol(g_cnt, zc, sizeof(zc));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_probe<<<1, 256, sm>>>(d_pk, d_pn, n, width, height, nIni, hX, N, M, d_out);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long p[8]; int c[8]; int outn;
        cudaMemcpyFromSymbol(p, g_prof, sizeof(p)); cudaMemcpyFromSymbol(c, g_cnt, sizeof(c));
        cudaMemcpy(&outn, d_out, 4, cudaMemcpyDeviceToHost);
        long long tot = 0; for (int i = 0; i < 8; i++) tot += p[i];
        printf("rep %d: n=%d N=%d nodes=%d kernel %.1f us, %lld cycles\n", rep, n, N, outn, ms * 1e3, tot);
        for (int i = 0; i < 8; i++) printf("  %-26s x%-3d %8lld cycles %5.1f%%\n", names[i], c[i], p[i], 100.0 * p[i] / tot);
    }
    return 0;
}
