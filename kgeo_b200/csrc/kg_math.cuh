// fp64 elementary functions for the per-sample hot loop (sm_100a): log, sincos, atan2, acos.
//
// Why not CUDA's libm: its log / sincos / atan2 / acos cost 75-135 SASS instructions per call in this kernel, and less than
// half of those are FP64 arithmetic -- every double-precision polynomial coefficient is an immediate that has to be built
// with two UMOV / IMAD.MOV instructions, and every call carries the subnormal / inf / NaN / huge-argument paths.  The
// theta-function evaluation of a point (kg_theta.cuh) needs 2 sincos, 1 acos and 3-4 of {log, atan2}: with libm that was
// 39 % of all issued instructions (ncu source page, profiles/).  Here the coefficients live in the constant bank (a DFMA
// takes a c[bank][offset] operand directly), arguments outside the plain range go to the libm function out of line (cold),
// and the divisions / square roots are the 1-ulp MUFU-seeded ones of kg_special.cuh.
//
// Accuracy (checked against libm on 1e7 arguments each by tests/hostsim, which compiles this same source for the host):
// <= 2 ulp.  Coefficients: research/gen_libm_coeffs.py (Chebyshev-node interpolation in 60 digits; polynomial error < 1e-17).
#pragma once
#include <string.h>

namespace kg {

// log: 7, sin: 7, cos: 6, atan: 8, asin: 13 coefficients
#define KG_MATH_COEFFS                                                                                                          \
    {6.66666666666666963e-01, 3.99999999998866151e-01, 2.85714286317643340e-01, 2.22222101992642113e-01,                        \
     1.81829566080634580e-01, 1.53295007542041783e-01, 1.46436286019097972e-01, /* log */                                       \
     -1.66666666666666657e-01, 8.33333333333333148e-03, -1.98412698412645666e-04, 2.75573192189547624e-06,                      \
     -2.50521061002559899e-08, 1.60585108753191948e-10, -7.58548648523593842e-13, /* sin */                                     \
     4.16666666666666644e-02, -1.38888888888872414e-03, 2.48015872985322604e-05, -2.75573171518588656e-07,                      \
     2.08761215346084536e-09, -1.13808681714499136e-11, /* cos */                                                               \
     -3.33333333333333315e-01, 1.99999999999970979e-01, -1.42857142841867985e-01, 1.11111108044739532e-01,                      \
     -9.09087880531204151e-02, 7.69067837978918423e-02, -6.61781995212533392e-02, 5.10868966915724310e-02, /* atan */           \
     1.66666666666666685e-01, 7.49999999999838851e-02, 4.46428571464460508e-02, 3.03819441312371992e-02,                        \
     2.23721732438322066e-02, 1.73523853934809472e-02, 1.39713253287287714e-02, 1.14780474725101455e-02,                        \
     1.03303702954044060e-02, 5.42423083980400102e-03, 1.74935694762482045e-02, -1.49998557838113486e-02,                       \
     2.88609743879985302e-02 /* asin */}
constexpr int KM_LOG = 0, KM_SIN = 7, KM_COS = 14, KM_ATAN = 20, KM_ASIN = 28, KM_COUNT = 41;
#if defined(__CUDACC__)
__constant__ double km_dev[KM_COUNT] = KG_MATH_COEFFS;
#endif
static const double km_host[KM_COUNT] = KG_MATH_COEFFS;
#if defined(__CUDA_ARCH__)
#define KM(i) km_dev[i]
#else
#define KM(i) km_host[i]
#endif

constexpr double kLn2Hi = 6.93147180369123816490e-01;   // 33 significant bits: e * kLn2Hi is exact
constexpr double kLn2Lo = 1.90821492927058770002e-10;
constexpr double kPio2Hi = 1.57079632679489655800e+00, kPio2Lo = 6.12323399573676603587e-17;
constexpr double kPiHi = 3.14159265358979311600e+00, kPiLo = 1.22464679914735317720e-16;
constexpr double kTwoOverPi = 6.36619772367581382433e-01;
constexpr double kTanPio16 = 0.19891236737965800691, kTan3Pio16 = 0.66817863791929891999;
constexpr double kTanPio8 = 0.41421356237309503;          // the double c; kAtanC = atan(c) for THIS double
constexpr double kAtanC = 0.39269908169872414;            // atan(0.41421356237309503) (= pi/8 - 1.1e-17)
constexpr double kPio4 = 0.78539816339744830962;

// out-of-line libm fallbacks for arguments outside the plain range (cold)
KG_NOINLINE double kg_log_cold(double x) { return log(x); }
KG_NOINLINE double kg_acos_cold(double x) { return acos(x); }
KG_NOINLINE double kg_atan2_cold(double y, double x) { return atan2(y, x); }
// (returned by value: with pointer out-parameters the CALLER's sine and cosine had to live on the stack on the hot path too --
// two 16-byte local stores and loads per sincos, four per point)
struct SinCos {
    double s, c;
};
KG_NOINLINE SinCos kg_sincos_cold(double v) {
    SinCos r;
    r.s = sin(v);
    r.c = cos(v);
    return r;
}

KG_HD int kg_hiint(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
#endif
}
KG_HD double kg_sethi(double x, int hi) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, __double2loint(x));
#else
    long long b;
    memcpy(&b, &x, 8);
    b = (b & 0xffffffffLL) | ((long long)hi << 32);
    memcpy(&x, &b, 8);
    return x;
#endif
}

// log(x): x = 2^e m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1) (|s| <= 0.1716)
KG_HD double flog(double x) {
    int hi = kg_hiint(x);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return kg_log_cold(x);   // zero, subnormal, negative, inf, NaN
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09f) {       // m >= ~sqrt(2): halve
        hi -= 0x00100000;
        e += 1;
    }
    const double m = kg_sethi(x, hi);
    const double s = fdiv(m - 1.0, m + 1.0);
    const double z = s * s;
    double p = KM(KM_LOG + 6);
#pragma unroll
    for (int i = 5; i >= 0; --i) p = fma(p, z, KM(KM_LOG + i));
    const double lm = fma(s * z, p, 2.0 * s);
    const double de = (double)e;
    return fma(de, kLn2Hi, fma(de, kLn2Lo, lm));
}

// (1/2) ln|a / b|, a and b finite and non-zero: the binary exponents of a and b are split off separately, so the only
// division is the s = (ma - mb)/(ma + mb) of the atanh series (ma - mb is exact)
KG_HD double half_log_ratio(double a, double b) {
    const int ha = kg_hiint(a) & 0x7fffffff, hb = kg_hiint(b) & 0x7fffffff;
    if ((unsigned)(ha - 0x00100000) >= 0x7fe00000u || (unsigned)(hb - 0x00100000) >= 0x7fe00000u)
        return 0.5 * kg_log_cold(fabs(a / b));                                  // zero, subnormal, inf, NaN
    int e = (ha >> 20) - (hb >> 20);
    double ma = kg_sethi(a, (ha & 0x000fffff) | 0x3ff00000), mb = kg_sethi(b, (hb & 0x000fffff) | 0x3ff00000);   // in [1, 2)
    if (ma > 1.4142135623730951 * mb) {            // ratio in (1/2, 2) -> [sqrt(1/2), sqrt(2)]
        mb += mb;
        e += 1;
    } else if (mb > 1.4142135623730951 * ma) {
        ma += ma;
        e -= 1;
    }
    const double s = fdiv(ma - mb, ma + mb);
    const double z = s * s;
    double p = KM(KM_LOG + 6);
#pragma unroll
    for (int i = 5; i >= 0; --i) p = fma(p, z, KM(KM_LOG + i));
    const double hl = fma(s * z, 0.5 * p, s);                                   // atanh(s) = (1/2) ln(ma/mb)
    const double de = 0.5 * (double)e;
    return fma(de, kLn2Hi, fma(de, kLn2Lo, hl));
}

// (1/2) ln|(X + Y)/(X - Y)| = atanh(Y/X) = atanh(X/Y): the Cauchy-principal-value logarithms of the third-kind integrals.
// One division on either path: the atanh series directly when the ratio is small, otherwise half_log_ratio.
KG_HD double atanh_ratio(double X, double Y) {
    const double ax = fabs(X), ay = fabs(Y);
    const bool swap = ay > ax;
    const double num = swap ? X : Y, den = swap ? Y : X;
    if (fabs(num) <= 0.17 * fabs(den)) {
        const double t = fdiv(num, den);
        const double z = t * t;
        double p = KM(KM_LOG + 6);
#pragma unroll
        for (int i = 5; i >= 0; --i) p = fma(p, z, KM(KM_LOG + i));
        return fma(t * z, 0.5 * p, t);
    }
    return half_log_ratio(X + Y, X - Y);
}

// sin and cos of v: Cody-Waite reduction by pi/2 (two FMAs: exact products), |v| < 1e6
KG_HD void fsincos(double v, double* sv, double* cv) {
    // |v| < 2^19 (and not NaN / inf): compared on the exponent field (integer pipe instead of an FP64-pipe DSETP)
    if ((unsigned)(kg_hiint(v) & 0x7fffffff) >= 0x41200000u) {
        const SinCos r = kg_sincos_cold(v);
        *sv = r.s;
        *cv = r.c;
        return;
    }
    const double fn = rint(v * kTwoOverPi);
    const int n = (int)fn;
    double r = fma(-fn, kPio2Hi, v);
    r = fma(-fn, kPio2Lo, r);
    const double z = r * r;
    double ps = KM(KM_SIN + 6), pc = KM(KM_COS + 5);
#pragma unroll
    for (int i = 5; i >= 0; --i) ps = fma(ps, z, KM(KM_SIN + i));
#pragma unroll
    for (int i = 4; i >= 0; --i) pc = fma(pc, z, KM(KM_COS + i));
    const double s = fma(r * z, ps, r);
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double a = (n & 1) ? c : s, b = (n & 1) ? s : c;       // quadrant 1, 3: swap
    // quadrants 2, 3 negate the sine, 1, 2 the cosine: flip the sign bit (the same value as -a / -b, without the
    // select on both register halves)
    *sv = kg_sethi(a, kg_hiint(a) ^ (int)(((unsigned)n & 2u) << 30));
    *cv = kg_sethi(b, kg_hiint(b) ^ (int)(((unsigned)(n + 1) & 2u) << 30));
}

// atan2(y, x), result in [-pi, pi]: one division; the octant angle is atan(c) + atan((a - c)/(1 + a c)) with
// c in {0, tan(pi/8), 1} chosen so that |t| <= tan(pi/16)
KG_HD double fatan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const bool swapped = ay > ax;
    const double mx = swapped ? ay : ax, mn = swapped ? ax : ay;
    // zeros, inf, NaN, extreme scales (either argument outside 2^-500 .. 2^500) go to libm; tested on the exponent fields
    if ((unsigned)((kg_hiint(ax) >> 20) - 523) > 1000u || (unsigned)((kg_hiint(ay) >> 20) - 523) > 1000u)
        return kg_atan2_cold(y, x);
    double c = 0.0, base = 0.0;
    if (mn > kTanPio16 * mx) {
        c = kTanPio8;
        base = kAtanC;
    }
    if (mn > kTan3Pio16 * mx) {
        c = 1.0;
        base = kPio4;
    }
    const double t = fdiv(fma(-c, mx, mn), fma(c, mn, mx));
    const double z = t * t;
    double p = KM(KM_ATAN + 7);
#pragma unroll
    for (int i = 6; i >= 0; --i) p = fma(p, z, KM(KM_ATAN + i));
    double r = base + fma(t * z, p, t);
    if (swapped) r = (kPio2Hi - r) + kPio2Lo;
    if (x < 0.0) r = (kPiHi - r) + kPiLo;
    return (y < 0.0) ? -r : r;
}

// acos(x), |x| <= 1
KG_HD double facos(double x) {
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return kg_acos_cold(x);
    if (ax <= 0.5) {
        const double z = x * x;
        double p = KM(KM_ASIN + 12);
#pragma unroll
        for (int i = 11; i >= 0; --i) p = fma(p, z, KM(KM_ASIN + i));
        return (kPio2Hi - fma(x * z, p, x)) + kPio2Lo;
    }
    const double z = 0.5 * (1.0 - ax);     // exact
    const double s = fsqrt(z);
    double p = KM(KM_ASIN + 12);
#pragma unroll
    for (int i = 11; i >= 0; --i) p = fma(p, z, KM(KM_ASIN + i));
    const double r = 2.0 * fma(s * z, p, s);
    return (x > 0.0) ? r : (kPiHi - r) + kPiLo;
}

}  // namespace kg
