// tests/native/fast_core_host.cpp -- compiles the product's packed FAST-9/16 arithmetic
// (orb-slam3_byzyh_b200/csrc/fast_core.h) for the host (lane emulation) so it can be unit-tested
// on CPU against the oracle's plain restatement.  A test of product logic, not a CPU path.
#include <cstdint>

#include "../../orb-slam3_byzyh_b200/csrc/fast_core.h"

// img: h x w bytes.  out[y*w+x] = margin max(best - sub, 0) for interior pixels, two pixels per call.
extern "C" void fast_core_margins(const uint8_t* img, int w, int h, int sub, uint8_t* out) {
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x + 1 < w - 3; x += 2) {
            const uint8_t* p = img + y * w + x;
            uint32_t e[16];
            const uint32_t c = ((uint32_t)p[0] | ((uint32_t)p[1] << 16)) + FC_BIAS2;
            for (int k = 0; k < 16; k++) {
                const uint8_t* q = p + dx[k] + dy[k] * w;
                e[k] = c - ((uint32_t)q[0] | ((uint32_t)q[1] << 16));
            }
            const uint32_t m = fc_margin2(e, (uint32_t)sub * 0x00010001u);
            out[y * w + x] = (uint8_t)(m & 0xFFFF);
            out[y * w + x + 1] = (uint8_t)(m >> 16);
        }
}
// the pair-sharing formulation (arcs k, k+1 share 8 ring positions)
extern "C" void fast_core_margins_pair(const uint8_t* img, int w, int h, int sub, uint8_t* out) {
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x + 1 < w - 3; x += 2) {
            const uint8_t* p = img + y * w + x;
            uint32_t e[16];
            const uint32_t c = ((uint32_t)p[0] | ((uint32_t)p[1] << 16)) + FC_BIAS2;
            for (int k = 0; k < 16; k++) {
                const uint8_t* q = p + dx[k] + dy[k] * w;
                e[k] = c - ((uint32_t)q[0] | ((uint32_t)q[1] << 16));
            }
            const uint32_t m = fc_margin2_pair(e, (uint32_t)sub * 0x00010001u);
            out[y * w + x] = (uint8_t)(m & 0xFFFF);
            out[y * w + x + 1] = (uint8_t)(m >> 16);
        }
}
// the raw-value formulation the kernel uses (no per-ring differences)
extern "C" void fast_core_margins_raw(const uint8_t* img, int w, int h, int sub, uint8_t* out) {
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x + 1 < w - 3; x += 2) {
            const uint8_t* p = img + y * w + x;
            uint32_t r[16];
            const uint32_t c = (uint32_t)p[0] | ((uint32_t)p[1] << 16);
            for (int k = 0; k < 16; k++) {
                const uint8_t* q = p + dx[k] + dy[k] * w;
                r[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 16);
            }
            const uint32_t m = fc_margin2_raw(r, c, (uint32_t)sub * 0x00010001u);
            out[y * w + x] = (uint8_t)(m & 0xFFFF);
            out[y * w + x + 1] = (uint8_t)(m >> 16);
        }
}
extern "C" void fast_core_margins_pair_raw(const uint8_t* img, int w, int h, int sub, uint8_t* out) {
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x + 1 < w - 3; x += 2) {
            const uint8_t* p = img + y * w + x;
            uint32_t r[16];
            const uint32_t c = (uint32_t)p[0] | ((uint32_t)p[1] << 16);
            for (int k = 0; k < 16; k++) {
                const uint8_t* q = p + dx[k] + dy[k] * w;
                r[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 16);
            }
            const uint32_t m = fc_margin2_pair_raw(r, c, (uint32_t)sub * 0x00010001u);
            out[y * w + x] = (uint8_t)(m & 0xFFFF);
            out[y * w + x + 1] = (uint8_t)(m >> 16);
        }
}
extern "C" void fast_core_margins_pair_raw_biased(const uint8_t* img, int w, int h, int sub, uint8_t* out) {
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x + 1 < w - 3; x += 2) {
            const uint8_t* p = img + y * w + x;
            uint32_t r[16];
            const uint32_t c = (uint32_t)p[0] | ((uint32_t)p[1] << 16);
            for (int k = 0; k < 16; k++) {
                const uint8_t* q = p + dx[k] + dy[k] * w;
                r[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 16);
            }
            const uint32_t m = fc_margin2_pair_raw_biased(r, c, (uint32_t)sub * 0x00010001u);
            out[y * w + x] = (uint8_t)(m & 0xFFFF);
            out[y * w + x + 1] = (uint8_t)(m >> 16);
        }
}
extern "C" int fast_core_best_scalar72(const uint8_t* p) { return fc_best_scalar<72>(p); }

// The dense early reject of k_fast_cells (fc_compass4), four pixels per call: out[y*w+x] = 1 where the pixel passes.
extern "C" void fast_core_compass(const uint8_t* img, int w, int h, int t, uint8_t* out) {
    const int u = t + 1;
    const uint32_t uLow = (uint32_t)(u & 0x7F) * 0x01010101u, uTop = (u & 0x80) ? 0xFFFFFFFFu : 0u;
    auto word = [&](int x, int y) {
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) v |= (uint32_t)img[y * w + x + i] << (8 * i);
        return v;
    };
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x + 3 < w - 3; x += 4) {
            const uint32_t f = fc_compass4<false>(word(x, y), word(x, y - 3), word(x, y + 3), word(x - 3, y), word(x + 3, y), uLow, uTop);
            if (u < 128 && f != fc_compass4<true>(word(x, y), word(x, y - 3), word(x, y + 3), word(x - 3, y), word(x + 3, y), uLow, uTop)) {
                out[0] = 255;   // the two variants disagree: make the test fail loudly
                return;
            }
            for (int i = 0; i < 4; i++) out[y * w + x + i] = (uint8_t)((f >> (8 * i + 7)) & 1);
        }
}
