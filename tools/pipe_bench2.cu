// tools/pipe_bench2.cu -- which pipe executes what on B200 (sm_100a)?  16 independent chains per
// thread, 32 warps per SM; reports warp-instructions per clock per SM sub-partition (SMSP).
// A 1:1 mix that runs faster than either opcode alone proves the two opcodes use different pipes.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define ITERS 1024
#define ILP 16
enum Op { V3 = 0, IM, H2, V3_IM, V3_H2, H2_IM, LOP, V3_LOP, POPC_, POPC_LOP, IADD_, V3_IADD, NOPS };
const char* names[] = {"VIMNMX3.U16x2", "IMAD", "HMNMX2", "VIMNMX3 + IMAD", "VIMNMX3 + HMNMX2", "HMNMX2 + IMAD", "LOP3",
                       "VIMNMX3 + LOP3", "POPC", "POPC + 3xLOP3", "IADD3", "VIMNMX3 + IADD3"};
__device__ __forceinline__ uint32_t hmx(uint32_t a, uint32_t b) { __half2 x = __hmax2(*(__half2*)&a, *(__half2*)&b); return *(uint32_t*)&x; }
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* clocks) {
    uint32_t a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed * (threadIdx.x + 1) + i * 977; b[i] = (seed >> 3) + i * 131 + threadIdx.x * 7; }
    uint32_t c = seed ^ 0x12345678u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const bool odd = i & 1;
            if (OP == V3) a[i] = __vimax3_u16x2(a[i], b[i], c);
            if (OP == IM) a[i] = a[i] * b[i] + c;
            if (OP == H2) a[i] = hmx(a[i], b[i]);
            if (OP == LOP) a[i] = (a[i] & b[i]) ^ c;
            if (OP == POPC_) a[i] = __popc(a[i] ^ c);
            if (OP == IADD_) a[i] = a[i] + b[i] + c;
            if (OP == V3_IM) a[i] = odd ? __vimax3_u16x2(a[i], b[i], c) : a[i] * b[i] + c;
            if (OP == V3_H2) a[i] = odd ? __vimax3_u16x2(a[i], b[i], c) : hmx(a[i], b[i]);
            if (OP == H2_IM) a[i] = odd ? hmx(a[i], b[i]) : a[i] * b[i] + c;
            if (OP == V3_LOP) a[i] = odd ? __vimax3_u16x2(a[i], b[i], c) : ((a[i] & b[i]) ^ c);
            if (OP == V3_IADD) a[i] = odd ? __vimax3_u16x2(a[i], b[i], c) : a[i] + b[i] + c;
            if (OP == POPC_LOP) { if ((i & 3) == 0) a[i] = __popc(a[i] ^ c); else a[i] = (a[i] & b[i]) ^ c; }
        }
        c += 0x00010001u;
    }
    long long t1 = clock64();
    uint32_t s = c;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}
template <int OP> void run(int sms, uint32_t* d_out, long long* d_clk) {
    const int ctas = sms * 4, threads = 256;
    for (int r = 0; r < 2; r++) k<OP><<<ctas, threads>>>(d_out, 12345u + r, d_clk);
    cudaDeviceSynchronize();
    std::vector<long long> clk(ctas);
    cudaMemcpy(clk.data(), d_clk, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto c : clk) avg += (double)c; avg /= ctas;
    // per SMSP: 8 warps x ITERS x ILP warp-instructions of the measured kind
    printf("%-22s %6.3f warp-instr/clk/SMSP\n", names[OP], 8.0 * ITERS * ILP / avg);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t* d_out; long long* d_clk;
    cudaMalloc(&d_out, 4ull * p.multiProcessorCount * 4 * 256); cudaMalloc(&d_clk, 8ull * p.multiProcessorCount * 4);
    const int n = p.multiProcessorCount;
    run<V3>(n, d_out, d_clk); run<IM>(n, d_out, d_clk); run<H2>(n, d_out, d_clk); run<LOP>(n, d_out, d_clk);
    run<IADD_>(n, d_out, d_clk); run<POPC_>(n, d_out, d_clk);
    run<V3_IM>(n, d_out, d_clk); run<V3_H2>(n, d_out, d_clk); run<H2_IM>(n, d_out, d_clk); run<V3_LOP>(n, d_out, d_clk);
    run<V3_IADD>(n, d_out, d_clk); run<POPC_LOP>(n, d_out, d_clk);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
}
