// sincosf_core.h -- sinf / cosf as the reference's host computes them, shared by k_describe and its host unit test
// (tests/test_sincosf_core.py compiles this header for the CPU and compares it with libm).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define SC_HD __device__ __forceinline__
#define SC_MUL(a, b) __dmul_rn((a), (b))
#define SC_ADD(a, b) __dadd_rn((a), (b))
#define SC_TRUNC(r) __double2int_rz(r)
#else
#define SC_HD static inline
#define SC_MUL(a, b) ((a) * (b))       /* host build: -ffp-contract=off */
#define SC_ADD(a, b) ((a) + (b))
#define SC_TRUNC(r) ((int)(r))
#endif

// sinf / cosf as the reference's host computes them.  `cos(angle)` / `sin(angle)` on a float (ORBextractor.cc:157) are the
// float overloads, i.e. glibc's sinf / cosf, and those are NOT correctly rounded: on 1.3 % of the angles one of the two
// differs by one ulp from the rounded fp64 value, which moves a rotated sample across a .5 boundary about once in a
// million descriptors (found by tools/parity_stress.py: one bit of one descriptor, sinf 0.517 ulp off).  glibc's routine
// (sysdeps/ieee754/flt-32/s_sincosf.h since 2.28: quadrant by a scaled fp64 multiply, x - n * pi/2 in fp64, a degree-7 /
// degree-8 fp64 polynomial, one rounding to float) is restated here operation by operation, without FMA; checked against
// libm on the host for 400 000 random angles (0 differences) and by the descriptor parity tests.  Also 25 fp64
// operations instead of the ~60 instructions of sincos().
SC_HD float glibc_sincosf_poly(double x, double x2, double sgn, int n) {
    if ((n & 1) == 0) {
        const double x3 = SC_MUL(x, x2);
        const double s1 = SC_ADD(0x1.1107605230bc4p-7, SC_MUL(x2, -0x1.994eb3774cf24p-13));
        const double x7 = SC_MUL(x3, x2);
        const double s = SC_ADD(x, SC_MUL(x3, -0x1.555545995a603p-3));
        return (float)SC_ADD(s, SC_MUL(x7, s1));
    }
    // the second table of the routine (n & 2) holds the negated cosine coefficients
    const double x4 = SC_MUL(x2, x2);
    const double c2 = SC_ADD(SC_MUL(sgn, -0x1.6c087e89a359dp-10), SC_MUL(x2, SC_MUL(sgn, 0x1.99343027bf8c3p-16)));
    const double c1 = SC_ADD(sgn, SC_MUL(x2, SC_MUL(sgn, -0x1.ffffffd0c621cp-2)));
    const double x6 = SC_MUL(x4, x2);
    const double c = SC_ADD(c1, SC_MUL(x4, SC_MUL(sgn, 0x1.55553e1068f19p-5)));
    return (float)SC_ADD(c, SC_MUL(x6, c2));
}
SC_HD void glibc_sincosf(float y, float* sn, float* cs) {
    const double x = (double)y;
    uint32_t bits;
    memcpy(&bits, &y, 4);
    const unsigned top = (bits >> 20) & 0x7ffu;
    if (top < 0x3f4u) {                       // |y| < 0.75: no reduction
        const double x2 = SC_MUL(x, x);
        const bool tiny = top < 0x398u;       // |y| < 2^-12
        *sn = tiny ? y : glibc_sincosf_poly(x, x2, 1.0, 0);
        *cs = tiny ? 1.0f : glibc_sincosf_poly(x, x2, 1.0, 1);
    } else if (top < 0x42fu) {                // |y| < 120: quadrant n = round(y * 2/pi), 24 fraction bits
        const double r = SC_MUL(x, 0x1.45F306DC9C883p+23);
        const int n = (SC_TRUNC(r) + 0x800000) >> 24;
        const double xr = SC_ADD(x, -SC_MUL((double)n, 0x1.921FB54442D18p0));
        const double sg = ((n + 1) & 2) ? -1.0 : 1.0;          // sign table {1, -1, -1, 1}[n & 3]
        const double tb = (n & 2) ? -1.0 : 1.0;
        const double xs = SC_MUL(xr, sg), x2 = SC_MUL(xr, xr);
        *sn = glibc_sincosf_poly(xs, x2, tb, n);
        *cs = glibc_sincosf_poly(xs, x2, tb, n ^ 1);
    } else {                                  // never an ORB angle: the large-argument path is not restated
#if defined(__CUDA_ARCH__)
        double sd, cd;
        sincos(x, &sd, &cd);
        *sn = (float)sd;
        *cs = (float)cd;
#else
        *sn = (float)__builtin_sin(x);
        *cs = (float)__builtin_cos(x);
#endif
    }
}
