// tools/pipe_bench4.cu -- issue rates on B200 (sm_100a) of the opcodes the round-2 kernels lean on: IDP2A / IDP4A (blur,
// resize), IMAD.HI (resize vertical pass), VABSDIFF4 + the SWAR byte compare (FAST dense reject), PRMT, SHF, two-input
// LOP3 / VIMNMX, LDS.32 and scattered LDS.U8 (FAST survivor scoring).  Same method as pipe_bench2.cu: 16 independent chains
// per thread, 32 warps per SM, warp-instructions per clock per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <vector>
#define ITERS 1024
#define ILP 16
enum Op { DP2A = 0, DP4A, MULHI, ABSD4, PRMT_, SHF_, LOP2, VMN2, LDS32, LDSU8, LDSU8R, NOPS };
const char* names[] = {"IDP2A", "IDP4A", "IMAD.HI", "VABSDIFF4", "PRMT", "SHF", "LOP3 (2 reg + imm)", "VIMNMX.U16x2 (2 in)",
                       "LDS.32 (conflict-free)", "LDS.U8 (same row, lane-strided)", "LDS.U8 (random in 2.3 KB)"};
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* clocks) {
    __shared__ uint32_t sm[8][640];
    for (int i = threadIdx.x; i < 8 * 640; i += 256) (&sm[0][0])[i] = i * 2654435761u + seed;
    __syncthreads();
    uint32_t a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed * (threadIdx.x + 1) + i * 977; b[i] = (seed >> 3) + i * 131 + threadIdx.x * 7; }
    uint32_t c = seed ^ 0x12345678u;
    const uint8_t* tile = reinterpret_cast<const uint8_t*>(sm[threadIdx.x >> 5]);
    const uint32_t* tw = sm[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == DP2A) a[i] = __dp2a_lo(b[i], c, a[i]);
            if (OP == DP4A) a[i] = __dp4a(b[i], c, a[i]);
            if (OP == MULHI) a[i] = __umulhi(a[i], b[i]) + 3u;   // IMAD.HI with the add folded in
            if (OP == ABSD4) a[i] = __vabsdiffu4(a[i], b[i]);
            if (OP == PRMT_) a[i] = __byte_perm(a[i], b[i], 0x5410 + (c & 0x1111));
            if (OP == SHF_) a[i] = __funnelshift_r(a[i], b[i], 8);
            if (OP == LOP2) a[i] = (a[i] | 0x80808080u) ^ b[i];
            if (OP == VMN2) a[i] = __vmaxu2(a[i], b[i]);
            if (OP == LDS32) a[i] = tw[(lane + i * 32 + (a[i] & 0x100)) & 511];
            if (OP == LDSU8) a[i] = tile[(lane * 3 + i * 48 + (a[i] & 0x400)) & 2047];
            if (OP == LDSU8R) a[i] = tile[(a[i] * 2654435761u >> 21) % 2300u];
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
    printf("%-34s %6.3f warp-instr/clk/SMSP (of the measured opcode; address / hash arithmetic of the LDS rows not counted)\n",
           names[OP], 8.0 * ITERS * ILP / avg);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t* d_out; long long* d_clk;
    cudaMalloc(&d_out, 4ull * p.multiProcessorCount * 4 * 256); cudaMalloc(&d_clk, 8ull * p.multiProcessorCount * 4);
    const int n = p.multiProcessorCount;
    run<DP2A>(n, d_out, d_clk); run<DP4A>(n, d_out, d_clk); run<MULHI>(n, d_out, d_clk); run<ABSD4>(n, d_out, d_clk);
    run<PRMT_>(n, d_out, d_clk); run<SHF_>(n, d_out, d_clk); run<LOP2>(n, d_out, d_clk); run<VMN2>(n, d_out, d_clk);
    run<LDS32>(n, d_out, d_clk); run<LDSU8>(n, d_out, d_clk); run<LDSU8R>(n, d_out, d_clk);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
}
