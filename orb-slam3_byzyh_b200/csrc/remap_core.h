// remap_core.h -- one sample of cv::remap(INTER_LINEAR, CV_32FC1 maps, BORDER_CONSTANT 0) on an 8-bit single-channel image,
// shared by intake.cu (orbfe_remap_linear) and pyramid.cu (rectification fused into level 0).  OpenCV 4.x arithmetic:
// coordinates rounded to 1/32 px (cvRound(x * 32), half to even), integer part kept as short, bilinear weights
// (32-fx)(32-fy)*32 ... summing to 2^15, result (sum + 2^14) >> 15, taps outside the source read as 0.
#pragma once
#include <stdint.h>

__device__ __forceinline__ int orbfe_remap_fix(float v) {
    // NaN / huge coordinates land far outside the source and read as border
    const float s = v * 32.0f;
    if (!(s > -1.0e9f)) return -(1 << 30);
    if (!(s < 1.0e9f)) return 1 << 30;
    return __float2int_rn(s);
}

__device__ __forceinline__ uint32_t orbfe_remap_sample(const uint8_t* __restrict__ src, size_t step, int srows, int scols,
                                                       float mx, float my) {
    const int sx = orbfe_remap_fix(mx), sy = orbfe_remap_fix(my);
    const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
    const int fx = sx & 31, fy = sy & 31;
    const bool x0 = ix >= 0 && ix < scols, x1 = ix + 1 >= 0 && ix + 1 < scols;
    const bool y0 = iy >= 0 && iy < srows, y1 = iy + 1 >= 0 && iy + 1 < srows;
    const uint8_t* r0 = src + (size_t)max(iy, 0) * step;
    const uint8_t* r1 = src + (size_t)max(iy + 1, 0) * step;
    const int p00 = (x0 && y0) ? r0[ix] : 0, p01 = (x1 && y0) ? r0[ix + 1] : 0;
    const int p10 = (x0 && y1) ? r1[ix] : 0, p11 = (x1 && y1) ? r1[ix + 1] : 0;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    return (uint32_t)((p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15);
}
