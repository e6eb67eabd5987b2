// Device-vs-host check of fast_core.h (debug tool).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../orb-slam3_byzyh_b200/csrc/fast_core.h"
constexpr int P = 72, ROWS = 22, NT = 4096;
__global__ void k(const uint8_t* tiles, int* best, int* quick) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NT) return;
    const uint8_t* p = tiles + (size_t)i * P * ROWS + 10 * P + 30;
    best[i] = fc_arc_best<P>(p);
    quick[i] = fc_may_be_corner<P>(p, 7) ? 1 : 0;
}
__global__ void k2(const uint8_t* tiles, int* best, int th) {   // same call pattern as k_fast_score
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NT) return;
    const uint8_t* p = tiles + (size_t)i * P * ROWS + 10 * P + 30;
    int b = 0;
    if (fc_may_be_corner<P>(p, th)) b = fc_arc_best<P>(p);
    best[i] = b > th ? b : 0;
}
int main() {
    std::vector<uint8_t> h((size_t)NT * P * ROWS);
    for (size_t i = 0; i < h.size(); i++) h[i] = (i / (P * ROWS)) % 3 == 0 ? rand() % 256 : 100 + rand() % 40;
    uint8_t* d; int *db, *dq, *db2;
    cudaMalloc(&d, h.size()); cudaMalloc(&db, NT * 4); cudaMalloc(&dq, NT * 4); cudaMalloc(&db2, NT * 4);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    k<<<NT / 128, 128>>>(d, db, dq);
    k2<<<NT / 128, 128>>>(d, db2, 7);
    std::vector<int> b(NT), q(NT), b2(NT);
    cudaMemcpy(b.data(), db, NT * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(q.data(), dq, NT * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(b2.data(), db2, NT * 4, cudaMemcpyDeviceToHost);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    int bad = 0, badq = 0, bad2 = 0;
    for (int i = 0; i < NT; i++) {
        const uint8_t* p = h.data() + (size_t)i * P * ROWS + 10 * P + 30;
        int hb = fc_arc_best<P>(p); int hq = fc_may_be_corner<P>(p, 7);
        int h2 = (hq && hb > 7) ? hb : 0;
        if (hb != b[i]) { if (bad < 5) printf("best mismatch %d: dev %d host %d\n", i, b[i], hb); bad++; }
        if (hq != q[i]) badq++;
        if (h2 != b2[i]) { if (bad2 < 5) printf("k2 mismatch %d: dev %d host %d\n", i, b2[i], h2); bad2++; }
    }
    printf("bad=%d badq=%d bad2=%d of %d\n", bad, badq, bad2, NT);
}
