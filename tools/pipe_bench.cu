// tools/pipe_bench.cu -- issue-rate microbenchmark of the integer / packed-16 instructions the hot
// kernels are built from (B200, sm_100a): ops per clock per SM for each opcode alone and in mixes.
// The roofline denominators for the INT-pipe-bound kernels (FAST score, Hamming) come from here.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define ITERS 2048
#define ILP 8

enum Op { VIMNMX3 = 0, VIMNMX2, HMNMX2, HFMA2, IMAD, POPC, LOP3, IADD, MIX_V3_H2, MIX_V3_IMAD, MIX_POPC_LOP3, MIX_V3_HFMA2R, PRMT, VIADD2, NOPS };
const char* names[] = {"VIMNMX3.U16x2", "VIMNMX.U16x2", "HMNMX2", "HFMA2", "IMAD", "POPC", "LOP3", "IADD3", "VIMNMX3+HMNMX2 (1:1)",
                       "VIMNMX3+IMAD (1:1)", "POPC+LOP3 (1:1)", "VIMNMX3+HFMA2.RELU (1:1)", "PRMT", "VIADD.16x2"};

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* clocks) {
    uint32_t a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed * (threadIdx.x + 1) + i * 977; b[i] = seed + i * 131 + threadIdx.x; }
    const uint32_t c = seed ^ 0x12345678u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == VIMNMX3) a[i] = __vimax3_u16x2(a[i], b[i], c);
            if (OP == VIMNMX2) a[i] = __vmaxu2(a[i], b[i]);
            if (OP == HMNMX2) { __half2 x = __hmax2(*(__half2*)&a[i], *(__half2*)&b[i]); a[i] = *(uint32_t*)&x; }
            if (OP == HFMA2) { __half2 x = __hfma2(*(__half2*)&a[i], *(__half2*)&b[i], *(__half2*)&c); a[i] = *(uint32_t*)&x; }
            if (OP == IMAD) a[i] = a[i] * b[i] + c;
            if (OP == POPC) a[i] = __popc(a[i]) + b[i];   // POPC + IADD
            if (OP == LOP3) a[i] = (a[i] ^ b[i]) | (a[i] & c);
            if (OP == IADD) a[i] = a[i] + b[i] + c;
            if (OP == PRMT) a[i] = __byte_perm(a[i], b[i], c);
            if (OP == VIADD2) a[i] = __vadd2(a[i], b[i]);
            if (OP == MIX_V3_H2) {
                if (i & 1) a[i] = __vimax3_u16x2(a[i], b[i], c);
                else { __half2 x = __hmax2(*(__half2*)&a[i], *(__half2*)&b[i]); a[i] = *(uint32_t*)&x; }
            }
            if (OP == MIX_V3_IMAD) {
                if (i & 1) a[i] = __vimax3_u16x2(a[i], b[i], c);
                else a[i] = a[i] * b[i] + c;
            }
            if (OP == MIX_POPC_LOP3) {
                if (i & 1) b[i] = __popc(a[i] ^ b[i - 1]);
                else a[i] = (a[i] ^ b[i]) | (a[i] & c);
            }
            if (OP == MIX_V3_HFMA2R) {
                if (i & 1) a[i] = __vimax3_u16x2(a[i], b[i], c);
                else { __half2 x = __hfma2_relu(*(__half2*)&a[i], *(__half2*)&b[i], *(__half2*)&c); a[i] = *(uint32_t*)&x; }
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, uint32_t* d_out, long long* d_clk) {
    const int ctas = sms * 4, threads = 256;   // 32 warps per SM
    k<OP><<<ctas, threads>>>(d_out, 12345u, d_clk);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<ctas, threads>>>(d_out, 999u, d_clk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> clk(ctas);
    cudaMemcpy(clk.data(), d_clk, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto c : clk) avg += (double)c; avg /= ctas;
    const double ops_per_sm = 4.0 * threads * (double)ITERS * ILP;   // lane-ops issued per SM
    printf("%-28s %8.1f lane-ops/clk/SM  (%6.3f ms, %.0f clk per CTA, eff. clock %.0f MHz)\n", names[OP],
           ops_per_sm / avg, ms, avg, avg / (ms * 1e3));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    uint32_t* d_out; long long* d_clk;
    cudaMalloc(&d_out, 4ull * p.multiProcessorCount * 4 * 256); cudaMalloc(&d_clk, 8ull * p.multiProcessorCount * 4);
    run<VIMNMX3>(p.multiProcessorCount, d_out, d_clk); run<VIMNMX2>(p.multiProcessorCount, d_out, d_clk);
    run<HMNMX2>(p.multiProcessorCount, d_out, d_clk); run<HFMA2>(p.multiProcessorCount, d_out, d_clk);
    run<IMAD>(p.multiProcessorCount, d_out, d_clk); run<POPC>(p.multiProcessorCount, d_out, d_clk);
    run<LOP3>(p.multiProcessorCount, d_out, d_clk); run<IADD>(p.multiProcessorCount, d_out, d_clk);
    run<PRMT>(p.multiProcessorCount, d_out, d_clk); run<VIADD2>(p.multiProcessorCount, d_out, d_clk);
    run<MIX_V3_H2>(p.multiProcessorCount, d_out, d_clk); run<MIX_V3_IMAD>(p.multiProcessorCount, d_out, d_clk);
    run<MIX_POPC_LOP3>(p.multiProcessorCount, d_out, d_clk); run<MIX_V3_HFMA2R>(p.multiProcessorCount, d_out, d_clk);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
