// sinf / cosf exactly as the host's glibc computes them (glibc >= 2.28, sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h:
// argument widened to double, fast range reduction by pi/2 for |x| < 120, degree-7 / degree-8 polynomials in double, one rounding
// to float).  The reference builds its rotation matrices with these two functions (jly_goicp.cpp:729-747, float overloads of
// <math.h>); the device-resident OuterBnB has to produce the same 9 floats per rotation cube, so the device evaluates the same
// operation sequence.  x86-64 glibc selects one of two builds of the functions at load time (ifunc): the FMA build contracts
// every multiply-add of the polynomials and of the range reduction into one fused operation, the SSE2 build rounds twice --
// `fma` selects which one is mirrored (engine: __builtin_cpu_supports("fma") && ("avx2"), glibc's own rule).
// Constants: __sincosf_table of the shipped libm (identical in every glibc since 2.28).
// Valid for |x| < 120 (rotation angles are <= sqrt(3) * pi); larger arguments return NaN (never reached).
#pragma once
#if defined(__CUDACC__)
#define LME_HD __host__ __device__ __forceinline__
#else
#define LME_HD static inline
#endif

namespace libm_exact {

LME_HD double lme_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
LME_HD double lme_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
LME_HD double lme_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
// a * b + c: fused (FMA build) or two roundings (SSE2 build)
LME_HD double lme_mad(bool fma, double a, double b, double c) { return fma ? lme_fma(a, b, c) : lme_add(lme_mul(a, b), c); }

// sinf_poly (sincosf.h): n even -> sine polynomial of x (x already multiplied by the quadrant sign), n odd -> cosine polynomial
LME_HD float lme_poly(bool fma, double x, double x2, bool neg, int n) {
    const double sg = neg ? -1.0 : 1.0;   // __sincosf_table[1] = the negated cosine coefficients (sine ones unchanged)
    if ((n & 1) == 0) {
        const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
        const double x3 = lme_mul(x, x2);
        const double s1 = lme_mad(fma, x2, S3, S2);
        const double x7 = lme_mul(x3, x2);
        const double s = lme_mad(fma, x3, S1, x);
        return (float)lme_mad(fma, x7, s1, s);
    }
    const double C0 = sg * 0x1p0, C1 = sg * -0x1.ffffffd0c621cp-2, C2 = sg * 0x1.55553e1068f19p-5, C3 = sg * -0x1.6c087e89a359dp-10, C4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = lme_mul(x2, x2);
    const double c2 = lme_mad(fma, x2, C4, C3);
    const double c1 = lme_mad(fma, x2, C1, C0);
    const double x6 = lme_mul(x4, x2);
    const double c = lme_mad(fma, x4, C2, c1);
    return (float)lme_mad(fma, x6, c2, c);
}
LME_HD unsigned lme_abstop12(float y) {
#if defined(__CUDA_ARCH__)
    return (__float_as_uint(y) >> 20) & 0x7ffu;
#else
    union { float f; unsigned u; } v; v.f = y; return (v.u >> 20) & 0x7ffu;
#endif
}
// reduce_fast: n = round(x * 2/pi), x - n * pi/2
LME_HD double lme_reduce(bool fma, double x, int* np) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    const double r = lme_mul(x, HPI_INV);
    const int n = ((int)r + 0x800000) >> 24;
    *np = n;
    return fma ? lme_fma(-(double)n, HPI, x) : lme_add(x, -lme_mul((double)n, HPI));
}
LME_HD float lme_sincos(bool fma, float y, int cosine) {
    const double x = (double)y;
    const unsigned top = lme_abstop12(y);
    if (top < 0x3f4u) {                       // |y| < pi/4
        if (top < 0x398u) return cosine ? 1.0f : y;   // |y| < 2^-12
        return lme_poly(fma, x, lme_mul(x, x), false, cosine);
    }
    if (top < 0x42fu) {                       // |y| < 120
        int n;
        const double xr = lme_reduce(fma, x, &n);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[n & 3] = {1, -1, -1, 1}
        return lme_poly(fma, lme_mul(xr, s), lme_mul(xr, xr), (n & 2) != 0, n ^ cosine);
    }
    return y - y + (0.0f / 0.0f) * 0.0f + (y - y) / (y - y);   // NaN: outside the range this engine ever asks for
}
LME_HD float sinf_glibc(bool fma, float y) { return lme_sincos(fma, y, 0); }
LME_HD float cosf_glibc(bool fma, float y) { return lme_sincos(fma, y, 1); }

}  // namespace libm_exact
