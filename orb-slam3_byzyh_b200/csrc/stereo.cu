// stereo.cu -- Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:1102-1358): rectified
// stereo association of left and right keypoints on the device-resident pyramids of the two
// extractors (the reference reads mpORBextractor{Left,Right}->mvImagePyramid, :1249-1275).
//
// One warp per LEFT keypoint.  The reference's per-row candidate table (:1132-1155) is only an
// index: right keypoint iR is a candidate of row (int)vL iff floor(yR-r) <= row <= ceil(yR+r),
// r = 2*mvScaleFactors[octave].  k_stereo_rows builds the same index per band of 16 image rows
// (count, scan, fill by one CTA per pair); the warp then evaluates the exact predicate plus the octave
// and disparity gates for the right keypoints of its band only, takes the Hamming minimum as a packed key
// (dist << 16 | iR: strict `<` in ascending-iR order == lowest iR among equal distances, start
// value TH_HIGH), then slides the 11x11 SAD window (+-5 px) with the 121 pixels spread over the
// lanes, fits the parabola and writes mvuRight / mvDepth.  A second single-CTA kernel applies
// the median-based outlier cut (:1343-1357) with a radix selection of the median instead of a sort.
#include <limits.h>

#include <algorithm>

#include "orbfe_internal.h"

namespace {

constexpr int TH_HIGH = 100, TH_LOW = 50;  // ORBmatcher.cc:36-37
constexpr int SB_H = 16;                   // image rows per candidate band

__device__ __forceinline__ void stereo_one(const OrbfeFrameGeom& g, const uint8_t* __restrict__ pyrL,
         const uint8_t* __restrict__ pyrR, const OrbfeKeyPoint* __restrict__ keysL,
         const uint32_t* __restrict__ descL, int N, const OrbfeKeyPoint* __restrict__ keysR,
         const uint32_t* __restrict__ descR, int Nr, float mbf, float mb, float* __restrict__ uRight,
         float* __restrict__ depth, int* __restrict__ sadOut, const int* __restrict__ bandStart,
         const int* __restrict__ bandList) {
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (iL >= N) return;
    if (lane == 0) { uRight[iL] = -1.0f; depth[iL] = -1.0f; sadOut[iL] = -1; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = g.lv[0].h;
    const OrbfeKeyPoint kpL = keysL[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows || levelL < 0 || levelL >= g.nlevels) return;
    const float minZ = mb, minD = 0.f, maxD = mbf / minZ;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) return;
    uint32_t dl[8];
    {
        const uint4* p = reinterpret_cast<const uint4*>(descL + 8 * (size_t)iL);
        *reinterpret_cast<uint4*>(dl) = p[0];
        *reinterpret_cast<uint4*>(dl + 4) = p[1];
    }
    uint32_t key = (uint32_t)TH_HIGH << 16;
    const int c0 = bandStart[row / SB_H], c1 = bandStart[row / SB_H + 1];
    for (int c = c0 + lane; c < c1; c += 32) {
        const int iR = bandList[c];
        const OrbfeKeyPoint kpR = keysR[iR];
        if (kpR.octave < 0 || kpR.octave >= g.nlevels) continue;
        const float r = 2.0f * g.lv[kpR.octave].scale;
        const int maxr = (int)ceilf(kpR.y + r), minr = (int)floorf(kpR.y - r);
        if (row < minr || row > maxr) continue;
        if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
        const float uR = kpR.x;
        if (!(uR >= minU && uR <= maxU)) continue;
        const uint4* p = reinterpret_cast<const uint4*>(descR + 8 * (size_t)iR);
        const uint4 a = p[0], b = p[1];
        const int dist = __popc(dl[0] ^ a.x) + __popc(dl[1] ^ a.y) + __popc(dl[2] ^ a.z) + __popc(dl[3] ^ a.w) +
                         __popc(dl[4] ^ b.x) + __popc(dl[5] ^ b.y) + __popc(dl[6] ^ b.z) + __popc(dl[7] ^ b.w);
        key = min(key, ((uint32_t)dist << 16) | (uint32_t)iR);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    const int bestDist = (int)(key >> 16);
    if (bestDist >= thOrbDist) return;
    const int bestIdxR = (int)(key & 0xFFFFu);

    // ---- sub-pixel refinement by sliding an 11x11 SAD window (:1233-1313) ----
    const float uR0 = keysR[bestIdxR].x;
    const float scaleFactor = g.lv[levelL].invScale;
    const float scaleduL = roundf(kpL.x * scaleFactor);
    const float scaledvL = roundf(kpL.y * scaleFactor);
    const float scaleduR0 = roundf(uR0 * scaleFactor);
    const int w = 5, Lw = 5;
    const OrbfeLevelGeom& G = g.lv[levelL];
    const float iniu = scaleduR0 + Lw - w;
    const float endu = scaleduR0 + Lw + w + 1;
    if (iniu < 0 || endu >= (float)G.w) return;
    const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w);
    // Coordinates are ROI-relative; the 19-px border keeps small excursions inside the buffer.
    if (y0 < -ORBFE_YOFF || y0 + 10 >= G.h + ORBFE_YOFF || xl0 < -ORBFE_EDGE || xl0 + 10 >= G.w + ORBFE_EDGE) return;
    const uint8_t* PL = pyrL + G.off + (size_t)ORBFE_YOFF * G.pitch + ORBFE_XOFF;
    const uint8_t* PR = pyrR + G.off + (size_t)ORBFE_YOFF * G.pitch + ORBFE_XOFF;
    int lv[4], yy[4], xx[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int p = lane + 32 * t;
        yy[t] = p / 11;
        xx[t] = p - yy[t] * 11;
        lv[t] = p < 121 ? (int)PL[(y0 + yy[t]) * G.pitch + xl0 + xx[t]] : 0;
    }
    int bestSad = INT_MAX, bestincR = 0;
    float vDists[11];
#pragma unroll
    for (int incR = -Lw; incR <= Lw; incR++) {
        const int xr0 = (int)(scaleduR0 + (float)incR - (float)w);
        int sad = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int p = lane + 32 * t;
            if (p < 121) sad += abs(lv[t] - (int)PR[(y0 + yy[t]) * G.pitch + xr0 + xx[t]]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sad += __shfl_xor_sync(0xffffffffu, sad, o);
        const float dist = (float)sad;  // cv::norm(IL, IR, NORM_L1) on CV_8U
        if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
        vDists[Lw + incR] = dist;
    }
    if (lane != 0) return;
    if (bestincR == -Lw || bestincR == Lw) return;
    float dist1 = 0.f, dist2 = 0.f, dist3 = 0.f;
#pragma unroll
    for (int k = 1; k < 10; k++)
        if (k == Lw + bestincR) { dist1 = vDists[k - 1]; dist2 = vDists[k]; dist3 = vDists[k + 1]; }
    const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = G.scale * ((float)scaleduR0 + (float)bestincR + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
            disparity = 0.01f;
            bestuR = (float)((double)uL - 0.01);
        }
        depth[iL] = mbf / disparity;
        uRight[iL] = bestuR;
        sadOut[iL] = bestSad;
    }
}

// Candidate index of the right keypoints of one pair (blockIdx.x): bandList[bandStart[b] .. bandStart[b + 1]) = the right
// keypoints whose row range [floor(y - r), ceil(y + r)] meets rows [16 b, 16 b + 15] (any order: the matcher's key
// carries the index).  At most `span` entries per keypoint, so span * capacity entries per pair.
__global__ void __launch_bounds__(256)
k_stereo_rows(const __grid_constant__ OrbfeFrameGeom g, const OrbfeKeyPoint* __restrict__ keysR, const int* __restrict__ nR,
              int NrHost, int capacity, int nb, int span, int* __restrict__ bandStart, int* __restrict__ bandList) {
    extern __shared__ int sb[];      // counts / cursors [nb], starts [nb + 1]
    int* cnt = sb;
    int* start = sb + nb;
    const size_t b = blockIdx.x;
    const int Nr = nR ? min(nR[b], capacity) : NrHost;
    const OrbfeKeyPoint* kr = keysR + b * (size_t)capacity;
    const int nRows = g.lv[0].h;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int pass = 0; pass < 2; pass++) {
        for (int iR = threadIdx.x; iR < Nr; iR += blockDim.x) {
            const OrbfeKeyPoint kp = kr[iR];
            if (kp.octave < 0 || kp.octave >= g.nlevels) continue;
            const float r = 2.0f * g.lv[kp.octave].scale;
            const int maxr = (int)ceilf(kp.y + r), minr = (int)floorf(kp.y - r);
            if (maxr < 0 || minr > nRows - 1) continue;
            const int b0 = max(minr, 0) / SB_H, b1 = min(min(maxr, nRows - 1) / SB_H, b0 + span - 1);
            for (int k = b0; k <= b1; k++) {
                const int pos = atomicAdd(&cnt[k], 1);
                if (pass == 1) bandList[b * (size_t)capacity * span + start[k] + pos] = iR;
            }
        }
        __syncthreads();
        if (pass == 0) {
            if (threadIdx.x == 0) {
                int acc = 0;
                for (int k = 0; k < nb; k++) { start[k] = acc; acc += cnt[k]; }
                start[nb] = acc;
            }
            __syncthreads();
            for (int i = threadIdx.x; i <= nb; i += blockDim.x) {
                bandStart[b * (size_t)(nb + 1) + i] = start[i];
                if (i < nb) cnt[i] = 0;
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256)
k_stereo(const __grid_constant__ OrbfeFrameGeom g, const uint8_t* __restrict__ pyrL,
         const uint8_t* __restrict__ pyrR, const OrbfeKeyPoint* __restrict__ keysL,
         const uint32_t* __restrict__ descL, int N, const OrbfeKeyPoint* __restrict__ keysR,
         const uint32_t* __restrict__ descR, int Nr, float mbf, float mb, float* __restrict__ uRight,
         float* __restrict__ depth, int* __restrict__ sadOut, const int* __restrict__ bandStart, const int* __restrict__ bandList) {
    stereo_one(g, pyrL, pyrR, keysL, descL, N, keysR, descR, Nr, mbf, mb, uRight, depth, sadOut, bandStart, bandList);
}

// The same for a batch of rectified pairs whose pyramids, keypoints and descriptors are resident in HBM (the output
// slabs of orbfe_extract_batch_device: `capacity` rows per frame, nL[b] / nR[b] valid): grid.y = pair.
__global__ void __launch_bounds__(256)
k_stereo_batch(const __grid_constant__ OrbfeFrameGeom g, const uint8_t* __restrict__ pyrL, const uint8_t* __restrict__ pyrR,
               const OrbfeKeyPoint* __restrict__ keysL, const uint32_t* __restrict__ descL, const int* __restrict__ nL,
               const OrbfeKeyPoint* __restrict__ keysR, const uint32_t* __restrict__ descR, const int* __restrict__ nR,
               int capacity, float mbf, float mb, float* __restrict__ uRight, float* __restrict__ depth, int* __restrict__ sadOut,
               int nb, int span, const int* __restrict__ bandStart, const int* __restrict__ bandList) {
    const size_t b = blockIdx.y, o = b * (size_t)capacity;
    stereo_one(g, pyrL + b * g.pyrStride, pyrR + b * g.pyrStride, keysL + o, descL + 8 * o, min(nL[b], capacity), keysR + o,
               descR + 8 * o, min(nR[b], capacity), mbf, mb, uRight + o, depth + o, sadOut + o, bandStart + b * (size_t)(nb + 1),
               bandList + o * span);
}

__device__ __forceinline__ void stereo_median_one(int N, float* __restrict__ uRight, float* __restrict__ depth, const int* __restrict__ sad);

// Median-based outlier cut (:1343-1357): median = sorted (SAD, iL) pairs [size/2].first.
__global__ void __launch_bounds__(1024)
k_stereo_median(int N, float* __restrict__ uRight, float* __restrict__ depth, const int* __restrict__ sad) {
    stereo_median_one(N, uRight, depth, sad);
}

__global__ void __launch_bounds__(1024)
k_stereo_median_batch(const int* __restrict__ nL, int capacity, float* __restrict__ uRight, float* __restrict__ depth,
                      const int* __restrict__ sad) {
    const size_t o = (size_t)blockIdx.x * capacity;
    stereo_median_one(min(nL[blockIdx.x], capacity), uRight + o, depth + o, sad + o);
}

// median = the element of rank n / 2 (0-based) of the n valid SADs, found by a two-level radix selection: a SAD of an
// 11 x 11 window of bytes is < 2^15, so 256 bins of the upper 8 bits, then 128 bins of the lower 7 inside the chosen bin.
__device__ __forceinline__ void stereo_median_one(int N, float* __restrict__ uRight, float* __restrict__ depth, const int* __restrict__ sad) {
    __shared__ int hist[256];
    __shared__ int s_bin, s_rank, s_median;
    const int nt = blockDim.x;
    for (int i = threadIdx.x; i < 256; i += nt) hist[i] = 0;
    if (threadIdx.x == 0) s_median = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += nt) {
        const int v = sad[i];
        if (v >= 0) atomicAdd(&hist[min(v >> 7, 255)], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < 256; k++) n += hist[k];
        s_bin = -1;
        if (n > 0) {
            int k = n / 2, acc = 0, bsel = 0;
            for (; bsel < 256; bsel++) {
                if (k < acc + hist[bsel]) break;
                acc += hist[bsel];
            }
            s_bin = bsel;
            s_rank = k - acc;
        }
    }
    __syncthreads();
    const int bin = s_bin;
    if (bin < 0) return;     // no match at all (the reference would index an empty vector)
    for (int i = threadIdx.x; i < 128; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += nt) {
        const int v = sad[i];
        if (v >= 0 && min(v >> 7, 255) == bin) atomicAdd(&hist[v & 127], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0, lo = 0;
        for (; lo < 128; lo++) {
            if (s_rank < acc + hist[lo]) break;
            acc += hist[lo];
        }
        s_median = (bin << 7) | lo;
    }
    __syncthreads();
    const float median = (float)s_median;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = threadIdx.x; i < N; i += nt) {
        const int v = sad[i];
        if (v >= 0 && !((float)v < thDist)) { uRight[i] = -1.f; depth[i] = -1.f; }
    }
}

}  // namespace

// bands of the candidate index and the most bands one right keypoint can meet (row range <= 2 * 2 * scale + 2 rows)
static void stereo_band_geometry(const OrbfeFrameGeom& g, int* nb, int* span) {
    float smax = 1.f;
    for (int l = 0; l < g.nlevels; l++) smax = g.lv[l].scale > smax ? g.lv[l].scale : smax;
    *nb = (g.lv[0].h + SB_H - 1) / SB_H;
    *span = (int)((4.f * smax + 2.f) / SB_H) + 2;
}

// ints of scratch the candidate index of B pairs of `capacity` keypoints needs
size_t orbfe_stereo_index_ints(const OrbfeFrameGeom& g, int B, int capacity) {
    int nb, span;
    stereo_band_geometry(g, &nb, &span);
    return (size_t)B * ((size_t)(nb + 1) + (size_t)std::max(capacity, 1) * span);
}

void orbfe_launch_stereo(const OrbfeFrameGeom& g, const uint8_t* pyrL, const uint8_t* pyrR,
                         const OrbfeKeyPoint* keysL, const uint32_t* descL, int N, const OrbfeKeyPoint* keysR,
                         const uint32_t* descR, int Nr, float mbf, float mb, float* uRight, float* depth,
                         int* sad, int* idx, cudaStream_t st) {
    int nb, span;
    stereo_band_geometry(g, &nb, &span);
    int* list = idx + (nb + 1);
    k_stereo_rows<<<1, 256, sizeof(int) * (2 * nb + 1), st>>>(g, keysR, nullptr, Nr, std::max(Nr, 1), nb, span, idx, list);
    k_stereo<<<(N + 7) / 8, 256, 0, st>>>(g, pyrL, pyrR, keysL, descL, N, keysR, descR, Nr, mbf, mb, uRight, depth, sad, idx, list);
    k_stereo_median<<<1, 256, 0, st>>>(N, uRight, depth, sad);
}

void orbfe_launch_stereo_batch(const OrbfeFrameGeom& g, const uint8_t* pyrL, const uint8_t* pyrR, int B,
                               const OrbfeKeyPoint* keysL, const uint32_t* descL, const int* nL, const OrbfeKeyPoint* keysR,
                               const uint32_t* descR, const int* nR, int capacity, float mbf, float mb, float* uRight,
                               float* depth, int* sad, int* idx, cudaStream_t st) {
    int nb, span;
    stereo_band_geometry(g, &nb, &span);
    int* list = idx + (size_t)B * (nb + 1);
    k_stereo_rows<<<B, 256, sizeof(int) * (2 * nb + 1), st>>>(g, keysR, nR, 0, capacity, nb, span, idx, list);
    k_stereo_batch<<<dim3((capacity + 7) / 8, B), 256, 0, st>>>(g, pyrL, pyrR, keysL, descL, nL, keysR, descR, nR, capacity, mbf,
                                                              mb, uRight, depth, sad, nb, span, idx, list);
    k_stereo_median_batch<<<B, 256, 0, st>>>(nL, capacity, uRight, depth, sad);
}
