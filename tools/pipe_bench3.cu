// tools/pipe_bench3.cu -- does operand reuse bound the FAST min/max network on B200?  VIMNMX3 reads three registers; the
// register file has two banks, so three distinct operands cost two fetch cycles unless the operand-reuse cache supplies
// some.  Variants (all: 16 values in registers, sliding 3-wise min AND max, 32 VIMNMX3 per round):
//   DISTINCT : a[i] = max3(a[i], b[i], d[i])                     three fresh operands per instruction
//   SLIDE    : m[k] = op3(r[k], r[k+1], r[k+2])                  natural operand order (slots shift every k)
//   STABLE   : same values, operands permuted so that two of the three stay in the SAME slot from k to k+1
//   PAIRED   : min3 and max3 of the same triple issued back to back (all three operands reusable)
#include <cstdio>
#include <cstdint>
#include <vector>
#define ITERS 512
enum { DISTINCT = 0, SLIDE, STABLE, PAIRED, NV };
const char* names[] = {"3 distinct operands", "sliding, natural order", "sliding, slot-stable", "min3/max3 paired"};
__device__ __forceinline__ uint32_t mn(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t mx(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* clocks) {
    uint32_t r[16], lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { r[i] = seed * (threadIdx.x + 1) + i * 977; lo[i] = r[i] ^ 0x5555u; hi[i] = r[i] + 3; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (V == DISTINCT) {
#pragma unroll
            for (int i = 0; i < 16; i++) { r[i] = mx(r[i], lo[i], hi[(i + 5) & 15]); lo[i] = mn(lo[i], hi[i], r[(i + 3) & 15]); }
        } else if (V == SLIDE) {
#pragma unroll
            for (int kk = 0; kk < 16; kk++) lo[kk] = mn(r[kk], r[(kk + 1) & 15], r[(kk + 2) & 15]);
#pragma unroll
            for (int kk = 0; kk < 16; kk++) hi[kk] = mx(r[kk], r[(kk + 1) & 15], r[(kk + 2) & 15]);
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = lo[i] + hi[(i + 1) & 15];
        } else if (V == STABLE) {
            // slot assignment: value r[j] always sits in slot j % 3, so consecutive triples share two slots
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                uint32_t s[3];
                s[kk % 3] = r[kk]; s[(kk + 1) % 3] = r[(kk + 1) & 15]; s[(kk + 2) % 3] = r[(kk + 2) & 15];
                lo[kk] = mn(s[0], s[1], s[2]);
            }
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                uint32_t s[3];
                s[kk % 3] = r[kk]; s[(kk + 1) % 3] = r[(kk + 1) & 15]; s[(kk + 2) % 3] = r[(kk + 2) & 15];
                hi[kk] = mx(s[0], s[1], s[2]);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = lo[i] + hi[(i + 1) & 15];
        } else {
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                lo[kk] = mn(r[kk], r[(kk + 1) & 15], r[(kk + 2) & 15]);
                hi[kk] = mx(r[kk], r[(kk + 1) & 15], r[(kk + 2) & 15]);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = lo[i] + hi[(i + 1) & 15];
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= r[i] ^ lo[i] ^ hi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}
template <int V> void run(int sms, uint32_t* d_out, long long* d_clk) {
    const int ctas = sms * 4;
    for (int rr = 0; rr < 2; rr++) k<V><<<ctas, 256>>>(d_out, 12345u + rr, d_clk);
    cudaDeviceSynchronize();
    std::vector<long long> clk(ctas);
    cudaMemcpy(clk.data(), d_clk, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto c : clk) avg += (double)c; avg /= ctas;
    const double instr = V == DISTINCT ? 32.0 : 48.0;   // VIMNMX3 (+16 IADD in the sliding variants) per round
    printf("%-26s %6.3f warp-instr/clk/SMSP (%.0f instr per round)\n", names[V], 8.0 * ITERS * instr / avg, instr);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t* d_out; long long* d_clk;
    cudaMalloc(&d_out, 4ull * p.multiProcessorCount * 4 * 256); cudaMalloc(&d_clk, 8ull * p.multiProcessorCount * 4);
    const int n = p.multiProcessorCount;
    run<DISTINCT>(n, d_out, d_clk); run<SLIDE>(n, d_out, d_clk); run<STABLE>(n, d_out, d_clk); run<PAIRED>(n, d_out, d_clk);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
}
