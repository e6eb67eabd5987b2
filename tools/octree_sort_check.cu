// tools/octree_sort_check.cu -- device check of the CTA-parallel std::sort emulation (csrc/octree_core.h:
// oc_std_sort_cta, warp-per-partition introsort loop + rank pass) against the real libstdc++ std::sort on the host.
// One CTA per array; many lengths, heavy ties, the depth-limit (heapsort) input; CTA sizes 128 and 256.
// Run by tests/test_gpu_octree_sort.py; exit code 0 = every permutation identical.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/octree_sort_check tools/octree_sort_check.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../orb-slam3_byzyh_b200/csrc/octree_core.h"

constexpr int CAP = 2048;   // longest array (a + out + queues = 36 KB of shared memory)

__global__ void k_sort(const uint64_t* in, uint64_t* outg, const int* offs) {
    __shared__ uint64_t a[CAP], out[CAP];
    __shared__ int q[4 * (CAP / 16 + 2)], cnt[3];
    const int o = offs[blockIdx.x], n = offs[blockIdx.x + 1] - o;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = in[o + i];
    __syncthreads();
    oc_std_sort_cta(a, out, n, q, cnt);
    for (int i = threadIdx.x; i < n; i += blockDim.x) outg[o + i] = out[i];
}

static bool key_less(const uint64_t& x, const uint64_t& y) { return (x >> 32) < (y >> 32); }

int main() {
    std::vector<uint64_t> in;
    std::vector<int> offs{0};
    srand(11);
    auto push = [&](const std::vector<uint64_t>& keys) {
        for (size_t i = 0; i < keys.size(); i++) in.push_back((keys[i] << 32) | (uint64_t)i);
        offs.push_back((int)in.size());
    };
    const int his[] = {1, 2, 4, 16, 100, 1000, 1 << 20};
    for (int n = 0; n <= 600; n += (n < 70 ? 1 : 7))
        for (int hi : his) {
            std::vector<uint64_t> k(n);
            for (auto& v : k) v = (uint64_t)(rand() % hi);
            push(k);
        }
    for (int rep = 0; rep < 40; rep++) {   // long arrays, sorted / reversed / organ-pipe shapes too
        const int n = 600 + rand() % (CAP - 600);
        std::vector<uint64_t> k(n);
        for (int i = 0; i < n; i++)
            k[i] = rep % 4 == 0 ? (uint64_t)i : rep % 4 == 1 ? (uint64_t)(n - i) : rep % 4 == 2 ? (uint64_t)(i < n / 2 ? i : n - i)
                                                                                              : (uint64_t)(rand() % 50);
        push(k);
    }
    {   // median-of-3 killer: drives introsort to its depth limit (heapsort fallback)
        const int n = CAP, half = n / 2;
        std::vector<uint64_t> k(n);
        for (int i = 0; i < half; i++) {
            k[i] = i % 2 == 0 ? (uint64_t)(i + 1) : (uint64_t)(half + i + (half % 2));
            k[half + i] = (uint64_t)(2 * (i + 1));
        }
        push(k);
    }
    const int nArr = (int)offs.size() - 1;
    std::vector<uint64_t> exp(in);
    for (int b = 0; b < nArr; b++) std::sort(exp.begin() + offs[b], exp.begin() + offs[b + 1], key_less);

    uint64_t *d_in, *d_out;
    int* d_offs;
    cudaMalloc(&d_in, in.size() * 8); cudaMalloc(&d_out, in.size() * 8); cudaMalloc(&d_offs, offs.size() * 4);
    cudaMemcpy(d_in, in.data(), in.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_offs, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice);
    int bad = 0;
    for (int nt : {128, 256}) {
        cudaMemset(d_out, 0xFF, in.size() * 8);
        k_sort<<<nArr, nt>>>(d_in, d_out, d_offs);
        std::vector<uint64_t> got(in.size());
        const cudaError_t e = cudaMemcpy(got.data(), d_out, in.size() * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        for (int b = 0; b < nArr; b++)
            if (!std::equal(exp.begin() + offs[b], exp.begin() + offs[b + 1], got.begin() + offs[b])) {
                if (bad < 10) printf("MISMATCH nt=%d array %d (n=%d)\n", nt, b, offs[b + 1] - offs[b]);
                bad++;
            }
    }
    printf("octree_sort_check: %d arrays x 2 CTA sizes, %d mismatches\n", nArr, bad);
    return bad ? 1 : 0;
}
