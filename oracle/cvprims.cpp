// oracle/cvprims.cpp -- TEST INFRASTRUCTURE ONLY. See cvprims.h for scope and provenance.
#include "cvprims.h"

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>

namespace cvp {

// ---------------------------------------------------------------------------------------
// resize, INTER_LINEAR, CV_8UC1.  OpenCV computes, per destination column dx,
//   fx = (float)((dx+0.5)*scale - 0.5), sx = floor(fx), fx -= sx, clamped at both ends,
// 11-bit integer weights cvRound(w*2048), a horizontal pass into int32 and a vertical pass
//   dst = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
// An exact 2x2 decimation is rerouted by OpenCV to its "area fast" path.
// ---------------------------------------------------------------------------------------
void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw,
                      int dh, size_t dstep) {
    if (dw == sw && dh == sh) {
        for (int y = 0; y < sh; y++) memcpy(dst + y * dstep, src + y * sstep, sw);
        return;
    }
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    const int iscale_x = cvRound(scale_x), iscale_y = cvRound(scale_y);
    const bool is_area_fast = std::abs(scale_x - iscale_x) < DBL_EPSILON &&
                              std::abs(scale_y - iscale_y) < DBL_EPSILON;
    if (is_area_fast && iscale_x == 2 && iscale_y == 2) {
        for (int dy = 0; dy < dh; dy++) {
            const uint8_t* s0 = src + (size_t)(2 * dy) * sstep;
            const uint8_t* s1 = s0 + sstep;
            uint8_t* d = dst + dy * dstep;
            for (int dx = 0; dx < dw; dx++)
                d[dx] = (uint8_t)((s0[2 * dx] + s0[2 * dx + 1] + s1[2 * dx] + s1[2 * dx + 1] + 2) >> 2);
        }
        return;
    }
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ia(2 * dw), ib(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ia[2 * dx] = (short)cvRound((1.f - fx) * 2048.f);
        ia[2 * dx + 1] = (short)cvRound(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloor(fy);
        fy -= sy;
        yofs[dy] = sy;
        ib[2 * dy] = (short)cvRound((1.f - fy) * 2048.f);
        ib[2 * dy + 1] = (short)cvRound(fy * 2048.f);
    }
    auto clipy = [sh](int y) { return y >= 0 ? (y < sh ? y : sh - 1) : 0; };
    std::vector<int> h0(dw), h1(dw);
    for (int dy = 0; dy < dh; dy++) {
        const uint8_t* s0 = src + (size_t)clipy(yofs[dy]) * sstep;
        const uint8_t* s1 = src + (size_t)clipy(yofs[dy] + 1) * sstep;
        for (int dx = 0; dx < dw; dx++) {
            const int sx = xofs[dx];
            const int sx1 = sx + 1 < sw ? sx + 1 : sx;  // weight is 0 when clamped
            h0[dx] = s0[sx] * ia[2 * dx] + s0[sx1] * ia[2 * dx + 1];
            h1[dx] = s1[sx] * ia[2 * dx] + s1[sx1] * ia[2 * dx + 1];
        }
        const int b0 = ib[2 * dy], b1 = ib[2 * dy + 1];
        uint8_t* d = dst + dy * dstep;
        for (int dx = 0; dx < dw; dx++)
            d[dx] = (uint8_t)((((b0 * (h0[dx] >> 4)) >> 16) + ((b1 * (h1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

void copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst,
                                 size_t dstep, int top, int bottom, int left, int right) {
    const int W = w + left + right, H = h + top + bottom;
    std::vector<int> xmap(W);
    for (int x = 0; x < W; x++) xmap[x] = reflect101(x - left, w);
    // src and dst may alias (dst interior == src, as in the reference's in-place call at
    // src/ORBextractor.cc:1712): interior rows first, then border rows from the interior.
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + y * sstep;
        uint8_t* d = dst + (size_t)(y + top) * dstep;
        std::vector<uint8_t> row(W);
        for (int x = 0; x < W; x++) row[x] = s[xmap[x]];
        memcpy(d, row.data(), W);
    }
    for (int y = 0; y < H; y++) {
        if (y >= top && y < top + h) continue;
        const int sy = reflect101(y - top, h);
        memcpy(dst + (size_t)y * dstep, dst + (size_t)(sy + top) * dstep, W);
    }
}

// ---------------------------------------------------------------------------------------
// FAST-9/16 (Bresenham ring r=3), OpenCV FAST_t<16> + cornerScore<16>.
// ---------------------------------------------------------------------------------------
void fast_ring_offsets(int step, int* ring16) {
    static const int off[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1},
                                   {2, -2}, {1, -3},  {0, -3},  {-1, -3}, {-2, -2}, {-3, -1},
                                   {-3, 0}, {-3, 1},  {-2, 2},  {-1, 3}};
    for (int k = 0; k < 16; k++) ring16[k] = off[k][0] + off[k][1] * step;
}

int fast_arc_best(const uint8_t* p, const int* ring16) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[ring16[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = 0;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) {
            mn = std::min(mn, d[s + k]);
            mx = std::max(mx, d[s + k]);
        }
        best = std::max(best, mn);   // centre brighter than the whole arc by at least mn
        best = std::max(best, -mx);  // centre darker than the whole arc by at least -mx
    }
    return best;
}

void fast9_16(const uint8_t* img, int w, int h, size_t step, int threshold, bool nms,
              std::vector<FastKP>& out) {
    out.clear();
    threshold = std::min(std::max(threshold, 0), 255);
    if (w < 7 || h < 7) return;
    int ring[16];
    fast_ring_offsets((int)step, ring);
    // score map: response (= best-1) for corners inside [3,w-3)x[3,h-3), 0 elsewhere.
    std::vector<uint8_t> sc((size_t)w * h, 0);
    std::vector<uint8_t> is((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = img + (size_t)y * step + x;
            // OpenCV's quick reject: a 9-arc contains one pixel of every opposite ring pair,
            // so each pair needs a member darker (bit 0) or brighter (bit 1) than the centre.
            const int v = p[0], lo = v - threshold, hi = v + threshold;
            int d = 3;
            for (int k = 0; k < 8 && d; k++) {
                const int a = p[ring[k]], b = p[ring[k + 8]];
                d &= ((a < lo) | ((a > hi) << 1)) | ((b < lo) | ((b > hi) << 1));
            }
            if (!d) continue;
            const int best = fast_arc_best(p, ring);
            if (best > threshold) {
                is[(size_t)y * w + x] = 1;
                sc[(size_t)y * w + x] = (uint8_t)(best - 1);
            }
        }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!is[(size_t)y * w + x]) continue;
            const int s = sc[(size_t)y * w + x];
            if (nms) {
                bool keep = true;
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        if (s <= sc[(size_t)(y + dy) * w + (x + dx)]) { keep = false; break; }
                    }
                if (!keep) continue;
            }
            out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------------------------------
// GaussianBlur 7x7 sigma=2 on u8: OpenCV >= 4 fixed-point path, 8.8 kernel
// [18 34 48 56 48 34 18], dst = (sum_y k_y * (sum_x k_x * src) + 32768) >> 16.
// ---------------------------------------------------------------------------------------
void gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst,
                          size_t dstep) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstep;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            for (int k = 0; k < 7; k++) acc += K[k] * s[reflect101(x + k - 3, w)];
            hbuf[(size_t)y * w + x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* d = dst + (size_t)y * dstep;
        const uint16_t* r[7];
        for (int k = 0; k < 7; k++) r[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; k++) acc += (uint32_t)K[k] * r[k][x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------
// fastAtan2: OpenCV's 7th-order odd minimax polynomial, evaluated in fp32 without FMA.
// ---------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

int hamming256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

void bf_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist) {
    for (int i = 0; i < nq; i++) {
        int b0 = 1 << 30, b1 = 1 << 30, i0 = -1, i1 = -1;
        for (int j = 0; j < nt; j++) {
            const int d = hamming256(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
            else if (d < b1) { b1 = d; i1 = j; }
        }
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
        dist[2 * i] = i0 < 0 ? -1 : b0; dist[2 * i + 1] = i1 < 0 ? -1 : b1;
    }
}

// ---------------------------------------------------------------------------------------
// cv::cvtColor to gray, 8-bit (OpenCV 4.x color_rgb: RY15 = 9798, GY15 = 19235, BY15 = 3735, shift 15).
// ---------------------------------------------------------------------------------------
void cvt_gray_u8(const uint8_t* src, int w, int h, size_t sstep, int channels, bool rgb, uint8_t* dst, size_t dstep) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = src + y * sstep + (size_t)x * channels;
            const int b = rgb ? p[2] : p[0], g = p[1], r = rgb ? p[0] : p[2];
            dst[y * dstep + x] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
        }
}

// ---------------------------------------------------------------------------------------
// cv::remap INTER_LINEAR, 8UC1, CV_32FC1 maps, BORDER_CONSTANT(0) (OpenCV imgwarp.cpp): coordinates are rounded to
// 1/32 px (INTER_BITS = 5) with cvRound, the integer part is kept as short, the four weights come from the
// fixed-point bilinear table (INTER_REMAP_COEF_BITS = 15; for bilinear the products are exact multiples of 32,
// so the table sums to 2^15 without the correction step), result = (sum + 2^14) >> 15.
// ---------------------------------------------------------------------------------------
void remap_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, int dw,
                     int dh, uint8_t* dst, size_t dstep) {
    auto fix = [](float v) {
        const float s = v * 32.0f;
        if (!(s > -1.0e9f)) return -(1 << 30);
        if (!(s < 1.0e9f)) return 1 << 30;
        return cvRound(s);
    };
    auto sat16 = [](int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); };
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            const int sx = fix(mapx[(size_t)y * dw + x]), sy = fix(mapy[(size_t)y * dw + x]);
            const int ix = sat16(sx >> 5), iy = sat16(sy >> 5), fx = sx & 31, fy = sy & 31;
            auto px = [&](int yy, int xx) { return (yy >= 0 && yy < sh && xx >= 0 && xx < sw) ? (int)src[yy * sstep + xx] : 0; };
            const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
            dst[y * dstep + x] = (uint8_t)((px(iy, ix) * w00 + px(iy, ix + 1) * w01 + px(iy + 1, ix) * w10 +
                                           px(iy + 1, ix + 1) * w11 + (1 << 14)) >> 15);
        }
}

// ---------------------------------------------------------------------------------------
// cv::undistortPoints with P = K and no rectification (Frame::UndistortKeyPoints, reference src/Frame.cc:1003-1051).
// ---------------------------------------------------------------------------------------
void undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, const double* dist, int ndist,
                      float* out) {
    double k[14] = {0};
    for (int i = 0; i < ndist && i < 14; i++) k[i] = dist[i];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; i++) {
        const double u = xy[2 * i], v = xy[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
}

}  // namespace cvp
