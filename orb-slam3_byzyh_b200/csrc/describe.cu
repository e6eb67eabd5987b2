// describe.cu -- the back half of ORBextractor::operator() (/root/reference/src/ORBextractor.cc):
//   k_blur     GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every level (:1629-1637)
//   k_layout   output ordering: level-major list order, scale to level-0 coordinates and the
//              vLappingArea front/back split (:1616-1678), as a block-wide scan per frame
//   k_describe IC_Angle (:91-138, computeOrientation :580-591) + steered BRIEF-256
//              (computeOrbDescriptor :150-203), one warp per retained keypoint
//
// OpenCV (not vendored by the reference) supplies GaussianBlur and fastAtan2; the arithmetic
// restated here is cv2 4.13's: the CV_8U blur is the 8.8 fixed-point kernel [18 34 48 56 48 34 18]
// with one rounding (sum + 32768) >> 16, and fastAtan2 is a 7th-order odd polynomial evaluated in
// fp32 WITHOUT fused multiply-add (oracle/cvprims.cpp; pinned by tests/test_oracle_cvprims.py).
// The 19-px reflect-101 border of the pyramid level is exactly the BORDER_REFLECT_101
// extrapolation GaussianBlur applies to the cloned ROI, so the blur reads the padded level.
#include <float.h>

#include "octree_core.h"  // OC_PK_*
#include "orbfe_internal.h"
#include "sincosf_core.h"

namespace {

__device__ const int8_t d_pattern[1024] = {
#include "orb_pattern.inc"
};

// ------------------------------------------------------------------------------------------
// k_blur: separable 7x7 fixed-point Gaussian, vertical pass first.
// A lane owns 4 adjacent pixels (one aligned 32-bit word of the padded level) and walks down a
// strip of BLUR_ROWS rows with a 7-row window in registers.  The vertical pass works on 16-bit
// lanes, two pixels per IMAD (column sums are <= 256*255 < 2^16, so nothing crosses a lane); the
// horizontal pass needs 24 bits: it stays on the packed pairs and runs as two-way dot products
// (IDP2A, 0.85 warp-instructions / clk / SMSP measured, tools/pipe_bench4.cu), its +-3 neighbours
// coming from the adjacent lanes by shuffle.  Warps overlap by one lane on each side (lanes 0 and 31 only feed halos), so
// a warp produces 120 columns.  No shared memory, every global access is a coalesced word.
constexpr int BLUR_ROWS = ORBFE_BLUR_TH, BLUR_WARPS = 4, BLUR_COLS = ORBFE_BLUR_TW;   // 32 rows, 120 columns per warp
static_assert(BLUR_COLS == 120, "one warp = 30 producing lanes x 4 pixels");

__global__ void __launch_bounds__(32 * BLUR_WARPS)
k_blur(const __grid_constant__ OrbfeFrameGeom g, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
    int l = 0;
    const int t = blockIdx.x;
    while (l + 1 < g.nlevels && t >= g.lv[l + 1].blurTileBase) l++;
    const OrbfeLevelGeom& L = g.lv[l];
    const int tl = t - L.blurTileBase;
    const int ty = tl / L.blurTilesX, tx = tl - ty * L.blurTilesX;
    const int lane = threadIdx.x & 31;
    const int y0 = (ty * BLUR_WARPS + (threadIdx.x >> 5)) * BLUR_ROWS;      // first output row (ROI)
    if (y0 >= L.h) return;
    const int x = BLUR_COLS * tx - 4 + 4 * lane;                             // ROI x of this lane's first pixel
    const size_t fo = (size_t)blockIdx.y * g.pyrStride + L.off;
    const int pw = L.pitch >> 2;
    const int wcol = min((ORBFE_XOFF + x) >> 2, pw - 1);                     // x >= -4: column >= 28
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pyr + fo) + wcol;
    uint32_t* dst = reinterpret_cast<uint32_t*>(blur + fo) + wcol;
    const int rmax = L.h + 2 * ORBFE_YOFF - 1;
    const bool store = lane >= 1 && lane <= 30 && x < L.w;

    uint32_t w01[7], w23[7];   // 7-row window, pixel pairs widened to 16-bit lanes
#pragma unroll
    for (int i = 0; i < BLUR_ROWS + 6; i++) {
        // padded row of input row (y0 - 3 + i)
        const uint32_t v = src[(size_t)min(ORBFE_YOFF + y0 - 3 + i, rmax) * pw];
        w01[i % 7] = __byte_perm(v, 0u, 0x4140);
        w23[i % 7] = __byte_perm(v, 0u, 0x4342);
        if (i >= 6) {
            // window rows (oldest..newest) = slots (i-6)%7 .. i%7
            const int r0 = (i - 6) % 7, r1 = (i - 5) % 7, r2 = (i - 4) % 7, r3 = (i - 3) % 7, r4 = (i - 2) % 7,
                      r5 = (i - 1) % 7, r6 = i % 7;
            const uint32_t V01 = 18u * (w01[r0] + w01[r6]) + 34u * (w01[r1] + w01[r5]) + 48u * (w01[r2] + w01[r4]) + 56u * w01[r3];
            const uint32_t V23 = 18u * (w23[r0] + w23[r6]) + 34u * (w23[r1] + w23[r5]) + 48u * (w23[r2] + w23[r4]) + 56u * w23[r3];
            // horizontal pass on the packed column sums: V[-4..7] as six 16-bit pairs (own, left and right neighbour
            // lane), every output = four two-way dot products (IDP2A: two 16-bit values x two 8-bit taps, 32-bit
            // accumulate) starting from the rounding constant -- 16 instructions for 4 pixels, no unpacking
            const uint32_t p01 = __shfl_up_sync(0xffffffffu, V01, 1), p23 = __shfl_up_sync(0xffffffffu, V23, 1);
            const uint32_t n01 = __shfl_down_sync(0xffffffffu, V01, 1), n23 = __shfl_down_sync(0xffffffffu, V23, 1);
            // tap pairs (low byte = tap of the pair's first value): kA = {(0,18), (34,48)}, kB = {(56,48), (34,18)},
            // kC = {(18,34), (48,56)}, kD = {(48,34), (18,0)} in the (lo, hi) halves
            const uint32_t kA = 0x30221200u, kB = 0x12223038u, kC = 0x38302212u, kD = 0x00122230u;
            uint32_t o0 = __dp2a_lo(p01, kA, 32768u);   // 18 V[-3]
            o0 = __dp2a_hi(p23, kA, o0);                // 34 V[-2] + 48 V[-1]
            o0 = __dp2a_lo(V01, kB, o0);                // 56 V[0] + 48 V[1]
            o0 = __dp2a_hi(V23, kB, o0);                // 34 V[2] + 18 V[3]
            uint32_t o1 = __dp2a_lo(p23, kC, 32768u);   // 18 V[-2] + 34 V[-1]
            o1 = __dp2a_hi(V01, kC, o1);                // 48 V[0] + 56 V[1]
            o1 = __dp2a_lo(V23, kD, o1);                // 48 V[2] + 34 V[3]
            o1 = __dp2a_hi(n01, kD, o1);                // 18 V[4]
            uint32_t o2 = __dp2a_lo(p23, kA, 32768u);   // 18 V[-1]
            o2 = __dp2a_hi(V01, kA, o2);                // 34 V[0] + 48 V[1]
            o2 = __dp2a_lo(V23, kB, o2);                // 56 V[2] + 48 V[3]
            o2 = __dp2a_hi(n01, kB, o2);                // 34 V[4] + 18 V[5]
            uint32_t o3 = __dp2a_lo(V01, kC, 32768u);   // 18 V[0] + 34 V[1]
            o3 = __dp2a_hi(V23, kC, o3);                // 48 V[2] + 56 V[3]
            o3 = __dp2a_lo(n01, kD, o3);                // 48 V[4] + 34 V[5]
            o3 = __dp2a_hi(n23, kD, o3);                // 18 V[6]
            const int y = y0 + i - 6;
            if (store && y < L.h) {
                // bytes 2 of each accumulator = (acc >> 16) & 255
                const uint32_t lo = __byte_perm(o0, o1, 0x0062), hi = __byte_perm(o2, o3, 0x0062);
                dst[(size_t)(ORBFE_YOFF + y) * pw] = __byte_perm(lo, hi, 0x5410);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_layout(const __grid_constant__ OrbfeFrameGeom g, const uint32_t* __restrict__ kp,
         const int* __restrict__ kpCount, OrbfeWork* __restrict__ work, int lap0, int lap1,
         OrbfeKeyPoint* __restrict__ outKps, int capacity, int* __restrict__ outN, int* __restrict__ outMono) {
    __shared__ int cnt[ORBFE_MAX_LEVELS];
    __shared__ int wsumS[8], wsumM[8];
    const int frame = blockIdx.x;
    if (threadIdx.x < g.nlevels) cnt[threadIdx.x] = kpCount[frame * g.nlevels + threadIdx.x];
    __syncthreads();
    int n = 0;
    for (int l = 0; l < g.nlevels; l++) n += cnt[l];
    const uint32_t* kpf = kp + (size_t)frame * g.kpCapFrame;
    OrbfeWork* wf = work + (size_t)frame * g.kpCapFrame;
    OrbfeKeyPoint* okp = outKps + (size_t)frame * capacity;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float flap0 = (float)lap0, flap1 = (float)lap1;
    int runS = 0, runM = 0;  // stereo / mono keypoints placed so far (slot order == reference order)
    for (int base = 0; base < g.kpCapFrame; base += 256) {
        const int s = base + threadIdx.x;
        bool valid = false, stereo = false;
        int l = 0, px = 0, py = 0, sc = 0;
        float fx = 0.f, fy = 0.f;
        if (s < g.kpCapFrame) {
            while (l + 1 < g.nlevels && s >= g.lv[l + 1].kpBase) l++;
            valid = (s - g.lv[l].kpBase) < cnt[l];
            if (valid) {
                const uint32_t p = kpf[s];
                px = OC_PK_X(p) + ORBFE_FAST_BORDER;   // :1184-1189
                py = OC_PK_Y(p) + ORBFE_FAST_BORDER;
                sc = OC_PK_S(p);
                fx = (float)px;
                fy = (float)py;
                if (l != 0) {                          // :1662-1665
                    fx = __fmul_rn(fx, g.lv[l].scale);
                    fy = __fmul_rn(fy, g.lv[l].scale);
                }
                stereo = fx >= flap0 && fx <= flap1;   // :1667
            }
        }
        const unsigned bS = __ballot_sync(0xffffffffu, valid && stereo);
        const unsigned bM = __ballot_sync(0xffffffffu, valid && !stereo);
        if (lane == 0) { wsumS[wid] = __popc(bS); wsumM[wid] = __popc(bM); }
        __syncthreads();
        int preS = 0, preM = 0, totS = 0, totM = 0;
#pragma unroll
        for (int w8 = 0; w8 < 8; w8++) {
            if (w8 < wid) { preS += wsumS[w8]; preM += wsumM[w8]; }
            totS += wsumS[w8];
            totM += wsumM[w8];
        }
        const unsigned lower = (1u << lane) - 1;
        if (s < g.kpCapFrame) {
            OrbfeWork wk;
            wk.x = (short)px; wk.y = (short)py; wk.level = (short)l; wk.pad = 0; wk.angle = -1.f;
            wk.dst = -1;
            if (valid) {
                const int dst = stereo ? n - 1 - (runS + preS + __popc(bS & lower))
                                       : runM + preM + __popc(bM & lower);
                if (dst < capacity) {
                    wk.dst = dst;
                    OrbfeKeyPoint k;
                    k.x = fx; k.y = fy; k.size = g.lv[l].kpsize; k.angle = -1.f;
                    k.response = (float)sc; k.octave = l; k.class_id = -1;
                    okp[dst] = k;
                }
            }
            wf[s] = wk;
        }
        runS += totS;
        runM += totM;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        outN[frame] = n;
        outMono[frame] = runM;
    }
}

// ------------------------------------------------------------------------------------------
// cv::fastAtan2 (degrees), fp32, no FMA contraction.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
    const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// One warp per keypoint, 8 keypoints per CTA, 32 registers (8 CTAs = 64 warps resident per SM).  The 256 test pairs are
// staged once per CTA as float4 (x0, y0, x1, y1), laid out [k][lane] so that lane `l` reads its k-th pair (pattern
// index 8 * l + k) with one conflict-free LDS.128 and no per-keypoint int8 -> float conversions.
__global__ void __launch_bounds__(256, 8)
k_describe(const __grid_constant__ OrbfeFrameGeom g, const uint8_t* __restrict__ pyr,
           const uint8_t* __restrict__ blur, const OrbfeWork* __restrict__ work,
           OrbfeKeyPoint* __restrict__ outKps, uint8_t* __restrict__ outDesc, int capacity) {
    __shared__ float4 patf[256];
    {
        const int p = 8 * (threadIdx.x & 31) + (threadIdx.x >> 5);
        const char4 q = reinterpret_cast<const char4*>(d_pattern)[p];
        patf[threadIdx.x] = make_float4((float)q.x, (float)q.y, (float)q.z, (float)q.w);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= g.kpCapFrame) return;
    const int frame = blockIdx.y;
    const OrbfeWork wk = work[(size_t)frame * g.kpCapFrame + s];
    if (wk.dst < 0) return;
    const OrbfeLevelGeom& L = g.lv[wk.level];
    const size_t co = (size_t)frame * g.pyrStride + L.off + (size_t)(ORBFE_YOFF + wk.y) * L.pitch + ORBFE_XOFF + wk.x;

    // ---- IC_Angle on the un-blurred level (:91-138) ----
    const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};  // ctor :542-570
    const uint8_t* c = pyr + co;
    int m10 = 0, m01 = 0;
    const int u = lane - ORBFE_HALF_PATCH;
    const int au = u < 0 ? -u : u;
    // rows +v and -v together (as :106-117 does): m10 = u * (sum of the column), m01 += v * (below - above)
    // Lane = column: a warp load is 31 adjacent bytes of one row (one cache line).  The two row pointers walk away from
    // the centre by one pitch per step (no per-row address arithmetic), rows outside the circular patch are read and
    // masked instead of branched around (they lie inside the level's 19-px border).
    const uint8_t* pp = c + u;
    const uint8_t* pm = pp;
    int sum = au <= 15 ? (int)pp[0] : 0;   // row v = 0 (lane 31 has au == 16: idle)
#pragma unroll
    for (int v = 1; v <= ORBFE_HALF_PATCH; v++) {
        pp += L.pitch;
        pm -= L.pitch;
        const int vp = pp[0], vm = pm[0];
        const bool in = au <= umax[v];
        sum += in ? vp + vm : 0;
        m01 += in ? v * (vp - vm) : 0;
    }
    m10 = u * sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered BRIEF on the blurred level (:150-203) ----
    const float factorPI = (float)(3.14159265358979323846 / 180.0);  // (float)(CV_PI/180.f), :141
    const float ang = __fmul_rn(angle, factorPI);
    float a, b;                         // a = cosf(ang), b = sinf(ang) of the reference's libm
    glibc_sincosf(ang, &b, &a);
    const uint8_t* bc = blur + co;
    unsigned val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float4 q = patf[32 * k + lane];
        const float x0 = q.x, y0 = q.y, x1 = q.z, y1 = q.w;
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = bc[r0 * L.pitch + c0], t1 = bc[r1 * L.pitch + c1];
        val |= (unsigned)(t0 < t1) << k;
    }
    outDesc[((size_t)frame * capacity + wk.dst) * 32 + lane] = (uint8_t)val;
    if (lane == 0) outKps[(size_t)frame * capacity + wk.dst].angle = angle;
}

}  // namespace

void orbfe_launch_blur(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, cudaStream_t st,
                       long long* launches) {
    k_blur<<<dim3(g.blurTiles, B), 32 * BLUR_WARPS, 0, st>>>(g, b.pyr, b.blur);
    ++*launches;
}

void orbfe_launch_layout(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B, int lap0, int lap1,
                         OrbfeKeyPoint* d_kps, int capacity, int* d_n, int* d_mono, cudaStream_t st,
                         long long* launches) {
    k_layout<<<B, 256, 0, st>>>(g, b.kp, b.kpCount, b.work, lap0, lap1, d_kps, capacity, d_n, d_mono);
    ++*launches;
}

void orbfe_launch_describe(const OrbfeFrameGeom& g, const OrbfeChunkBufs& b, int B,
                           OrbfeKeyPoint* d_kps, uint8_t* d_desc, int capacity, cudaStream_t st,
                           long long* launches) {
    k_describe<<<dim3((g.kpCapFrame + 7) / 8, B), 256, 0, st>>>(g, b.pyr, b.blur, b.work, d_kps, d_desc, capacity);
    ++*launches;
}
