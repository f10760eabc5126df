// fp64 special functions for the analytic Kerr ray tracer (sm_100a).
//
// Replaces the reference's L1 layer: scipy.special.{elliprf,elliprd,elliprj,elliprc,ellipkinc,ellipeinc,
// ellipj,ellipk} and kgeo/scipy_ellip_binding.py:87-237 (ellip_pi_arr).
//
// Design notes
//  * Everything is a __host__ __device__ inline function so the same source is (a) inlined into the fused
//    kernels and (b) compilable by the host compiler for the developer-only harness in tests/hostsim.
//    The product (libkgeo_b200.so) only ever launches the __global__ kernels.
//  * Carlson symmetric integrals by duplication with the degree-7 series of DLMF 19.36.1-2, so ~4 duplication
//    steps reach 1e-16.  RC(1,1+e) inside RJ is evaluated in closed form (atan / log) on the first step and by
//    its alternating series afterwards; 1+e is formed cancellation-free as 2*sqrt(p)*(p+lambda)/d.
//  * `carlson_bundle` runs ONE shared (x,y,1) duplication track for RF, RD and up to three RJ's: this is the hot
//    loop of the fused track kernel (one Legendre E and up to three Legendre Pi's share every sqrt).
//  * Jacobi sn,cn,dn by the descending Landen/AGM ladder (Bulirsch); K(m) falls out of the ladder for free, so
//    amplitudes are never formed: F(am u) = u exactly, and period reduction is done on u with K.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define KG_HD __host__ __device__ __forceinline__
#define KG_NOINLINE __host__ __device__ __noinline__
#else
#define KG_HD inline
#define KG_NOINLINE inline
#endif

namespace kg {

constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kHalfPi = 1.570796326794896619231321691639751442;
#define KG_INF (__builtin_huge_val())
constexpr double kInf = KG_INF;
constexpr double kSeriesTol = 0.0075;
// the shared-track bundle of the fused kernels stops one half step earlier on average: truncation error ~ 3 r^8 = 8e-15
// relative (measured: no change of the parity errors on any fixture; at 0.03 they grow 6x)
constexpr double kBundleTol = 0.015;  // |X|,|Y|,|Z|,|P| bound at which the degree-7 series is exact to 1e-16
constexpr int kMaxDup = 40;

KG_HD double sq(double x) { return x * x; }
KG_HD double cube(double x) { return x * x * x; }
KG_HD double max3(double a, double b, double c) { return fmax(a, fmax(b, c)); }
KG_HD double kg_nan() { return kInf - kInf; }

// ---- hot-loop arithmetic ---------------------------------------------------------------------------------------------
// The IEEE-rounded sqrt / division that nvcc emits cost ~45 / ~32 SASS instructions each (seed + Newton + fix-up paths),
// and the per-sample evaluation needs ~60 square roots and ~30 divisions.  The fused kernels use these instead: the
// MUFU seed (20 good bits), ONE third-order Newton step (error -> e^3 ~ 1e-18) and one residual correction, i.e. 6-9
// DFMA-class instructions, result within 1 ulp, no fix-up branches and no range checks: the Carlson / Landen recurrences
// only produce arguments in (1e-290, 1e290) or exact zeros (handled in fsqrt).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double kg_rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double kg_rsqrt_seed(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
// 1/sqrt(x), x in (1e-290, 1e290)
__device__ __forceinline__ double frsqrt(double x) {
    double y = kg_rsqrt_seed(x);
    double t = fma(-(x * y), y, 1.0);                 // 1 - x y^2
    double u = fma(0.375, t, 0.5) * t;                // t/2 + 3 t^2/8
    y = fma(y, u, y);
    t = fma(-(x * y), y, 1.0);                        // residual correction
    return fma(0.5 * y, t, y);
}
// sqrt(x) for x in (1e-300, 1e290) (callers clamp exact zeros to 1e-300)
__device__ __forceinline__ double fsqrt(double x) {
    double y = kg_rsqrt_seed(x);
    double t = fma(-(x * y), y, 1.0);
    double u = fma(0.375, t, 0.5) * t;
    y = fma(y, u, y);
    double g = x * y;
    double d = fma(-g, g, x);
    return fma(0.5 * y, d, g);
}
// a / b for |b| in (1e-290, 1e290) (no zero / infinity handling: the hot loops never divide by them)
__device__ __forceinline__ double fdiv(double a, double b) {
    double r = kg_rcp_seed(b);
    double t = fma(-b, r, 1.0);
    double u = fma(t, t, t);                          // t + t^2
    r = fma(r, u, r);
    double q = a * r;
    double rem = fma(-b, q, a);
    return fma(rem, r, q);
}
__device__ __forceinline__ double frcp(double b) { return fdiv(1.0, b); }
#else
KG_HD double frsqrt(double x) { return 1.0 / sqrt(x); }
KG_HD double fsqrt(double x) { return sqrt(x); }
KG_HD double fdiv(double a, double b) { return a / b; }
KG_HD double frcp(double b) { return 1.0 / b; }
#endif

// Series coefficients live in the constant bank on the device: an FP64 literal costs two UMOV instructions per use,
// a constant-bank operand costs none.
#define KG_SERIES_COEFFS                                                                                             \
    {1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0, /* 0-7  rc1p   */ \
     1.0 / 10.0, 1.0 / 14.0, 1.0 / 24.0, 3.0 / 44.0, 5.0 / 208.0, 3.0 / 104.0, 1.0 / 16.0,       /* 8-14 RF     */ \
     3.0 / 14.0, 1.0 / 6.0, 9.0 / 88.0, 3.0 / 22.0, 9.0 / 52.0, 3.0 / 26.0, 1.0 / 16.0, 3.0 / 40.0, 3.0 / 20.0,     \
     45.0 / 272.0, 9.0 / 68.0,                                                                   /* 15-25 RJ/RD */ \
     0.0075, 0.375, 0.25, 0.2, 1.0 / 3.0, 0.5}                                                   /* 26-31 misc  */
#if defined(__CUDACC__)
__constant__ double kc_dev[32] = KG_SERIES_COEFFS;
#endif
static const double kc_host[32] = KG_SERIES_COEFFS;
#if defined(__CUDA_ARCH__)
#define KC(i) kc_dev[i]
#else
#define KC(i) kc_host[i]
#endif

}  // namespace kg
#include "kg_math.cuh"
namespace kg {

// RC(1, b) with e = b - 1 given separately (both cancellation-free), e > -1:
//   atan(sqrt e)/sqrt e   (e>0),   atanh(sqrt -e)/sqrt -e = log((1+s)/sqrt(b))/s, s = sqrt(-e)   (e<0)
// closed forms (out of line: one copy serves every Pi track of the fused kernels)
KG_NOINLINE double rc1p_closed(double e, double b) {
    if (e > 0.0) {
        double s = sqrt(e);
        return atan(s) / s;
    }
    double s = sqrt(-e);
    return log((1.0 + s) / sqrt(b)) / s;
}
// same with b = 1 + e = 2 sp (p + lam) id formed only on the (rare) closed-form path
KG_HD double rc1p(double e, double b);
KG_HD double rc1p_series(double e);
KG_HD double rc1p_lazy(double e, double sp, double ppl, double id) {
    if (fabs(e) < 0.02) return rc1p_series(e);
    return rc1p_closed(e, 2.0 * sp * ppl * id);
}
KG_HD double rc1p_series(double e) {
    double me = -e;
    double h = KC(0);
    h = fma(h, me, KC(1));
    h = fma(h, me, KC(2));
    h = fma(h, me, KC(3));
    h = fma(h, me, KC(4));
    h = fma(h, me, KC(5));
    h = fma(h, me, KC(6));
    h = fma(h, me, KC(7));
    return fma(h, me, 1.0);
}
KG_HD double rc1p(double e, double b) {
    if (fabs(e) < 0.02) {
        // sum_k (-e)^k / (2k+1), k = 0..8   (0.02^9/19 < 3e-17), Horner in (-e)
        double me = -e;
        double h = KC(0);
        h = fma(h, me, KC(1));
        h = fma(h, me, KC(2));
        h = fma(h, me, KC(3));
        h = fma(h, me, KC(4));
        h = fma(h, me, KC(5));
        h = fma(h, me, KC(6));
        h = fma(h, me, KC(7));
        h = fma(h, me, 1.0);
        return h;
    }
    return rc1p_closed(e, b);
}

// ---- degree-7 tails (DLMF 19.36.1, 19.36.2) --------------------------------------------------------------
KG_HD double rf_series(double E2, double E3) {
    return 1.0 - E2 * KC(8) + E3 * KC(9) + E2 * E2 * KC(10) - E2 * E3 * KC(11) - E2 * E2 * E2 * KC(12) + E3 * E3 * KC(13) +
           E2 * E2 * E3 * KC(14);
}
KG_HD double rj_series(double E2, double E3, double E4, double E5) {
    return 1.0 - E2 * KC(15) + E3 * KC(16) + E2 * E2 * KC(17) - E4 * KC(18) - E2 * E3 * KC(19) + E5 * KC(20) -
           E2 * E2 * E2 * KC(21) + E3 * E3 * KC(22) + E2 * E4 * KC(23) + E2 * E2 * E3 * KC(24) - (E3 * E4 + E2 * E5) * KC(25);
}

// ---- stand-alone Carlson integrals (general arguments; exported for the element-wise C-ABI) ---------------
KG_HD double carlson_rf(double x, double y, double z) {
    if (!(x >= 0.0) || !(y >= 0.0) || !(z >= 0.0)) return kg_nan();
    double A = (x + y + z) * (1.0 / 3.0);
    for (int n = 0; n < kMaxDup; ++n) {
        if (max3(fabs(A - x), fabs(A - y), fabs(A - z)) < kSeriesTol * A) break;
        double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
        double lam = fma(sx, sy, fma(sy, sz, sz * sx));
        A = (A + lam) * 0.25;
        x = (x + lam) * 0.25;
        y = (y + lam) * 0.25;
        z = (z + lam) * 0.25;
    }
    double iA = 1.0 / A;
    double X = (A - x) * iA, Y = (A - y) * iA, Z = -(X + Y);
    double E2 = X * Y - Z * Z, E3 = X * Y * Z;
    return rf_series(E2, E3) * sqrt(iA);
}

KG_HD double carlson_rc(double x, double y) {
    // RC(x,y) = RF(x,y,y); closed forms; Cauchy principal value for y<0 (DLMF 19.2.20)
    if (!(x >= 0.0) || y == 0.0) return (y == 0.0) ? kInf : kg_nan();
    if (y < 0.0) return sqrt(x / (x - y)) * carlson_rc(x - y, -y);
    if (x == 0.0) return kHalfPi / sqrt(y);
    return rc1p((y - x) / x, y / x) / sqrt(x);
}

KG_HD double carlson_rd(double x, double y, double z) {
    if (!(x >= 0.0) || !(y >= 0.0) || !(z > 0.0)) return kg_nan();
    double A = (x + y + 3.0 * z) * 0.2;
    double f = 1.0, sum = 0.0;
    for (int n = 0; n < kMaxDup; ++n) {
        if (max3(fabs(A - x), fabs(A - y), fabs(A - z)) < kSeriesTol * A) break;
        double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
        double lam = fma(sx, sy, fma(sy, sz, sz * sx));
        sum += f / (sz * (z + lam));
        A = (A + lam) * 0.25;
        x = (x + lam) * 0.25;
        y = (y + lam) * 0.25;
        z = (z + lam) * 0.25;
        f *= 0.25;
    }
    double iA = 1.0 / A;
    double X = (A - x) * iA, Y = (A - y) * iA, Z = -(X + Y) * (1.0 / 3.0);
    double XY = X * Y, Z2 = Z * Z;
    double E2 = XY - 6.0 * Z2, E3 = (3.0 * XY - 8.0 * Z2) * Z, E4 = 3.0 * (XY - Z2) * Z2, E5 = XY * Z2 * Z;
    return f * rj_series(E2, E3, E4, E5) * iA * sqrt(iA) + 3.0 * sum;
}

KG_HD double carlson_rj(double x, double y, double z, double p) {
    if (!(x >= 0.0) || !(y >= 0.0) || !(z >= 0.0) || !(p > 0.0)) return kg_nan();
    double A = (x + y + z + 2.0 * p) * 0.2;
    double f = 1.0, sum = 0.0;
    for (int n = 0; n < kMaxDup; ++n) {
        if (fmax(max3(fabs(A - x), fabs(A - y), fabs(A - z)), fabs(A - p)) < kSeriesTol * A) break;
        double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z), sp = sqrt(p);
        double lam = fma(sx, sy, fma(sy, sz, sz * sx));
        double id = 1.0 / ((sp + sx) * (sp + sy) * (sp + sz));
        double e = (sp - sx) * (sp - sy) * (sp - sz) * id;
        double b = 2.0 * sp * (p + lam) * id;                       // = 1 + e without cancellation
        sum += f * id * rc1p(e, b);
        A = (A + lam) * 0.25;
        x = (x + lam) * 0.25;
        y = (y + lam) * 0.25;
        z = (z + lam) * 0.25;
        p = (p + lam) * 0.25;
        f *= 0.25;
    }
    double iA = 1.0 / A;
    double X = (A - x) * iA, Y = (A - y) * iA, Z = (A - z) * iA, P = -(X + Y + Z) * 0.5;
    double XYZ = X * Y * Z, P2 = P * P;
    double E2 = X * Y + X * Z + Y * Z - 3.0 * P2;
    double E3 = XYZ + 2.0 * E2 * P + 4.0 * P2 * P;
    double E4 = (2.0 * XYZ + E2 * P + 3.0 * P2 * P) * P;
    double E5 = XYZ * P2;
    return f * rj_series(E2, E3, E4, E5) * iA * sqrt(iA) + 6.0 * sum;
}

// ---- shared-track bundle: RF(x,y,1), RD(x,y,1), RJ(x,y,1,p_i), i < NP ---------------------------------------
// All integrals share x_n, y_n, z_n and lambda_n; each has its own mean A (RF: (x+y+z)/3, RD: (x+y+3z)/5,
// RJ_i: (x+y+z+2p_i)/5) which obeys the same recurrence A <- (A+lambda)/4.
template <int NP, bool WANT_RF, bool WANT_RD>
KG_HD void carlson_bundle(double x, double y, const double* p_in, double* rf, double* rd, double* rj) {
    // cos^2 of an amplitude at a turning point is an exact 0: RF/RD/RJ(1e-300, y, 1, p) equal the x = 0 values to ~1e-150
    // relative, and the clamp keeps the seed-based square root free of special cases; 0*x keeps a NaN a NaN.
    x = fma(0.0, x, fmax(x, 1.0e-300));
    y = fma(0.0, y, fmax(y, 1.0e-300));
    double z = 1.0;
    double p[NP > 0 ? NP : 1], Aj[NP > 0 ? NP : 1], sj[NP > 0 ? NP : 1], dl[NP > 0 ? NP : 1];
    double Af = (x + y + z) * (1.0 / 3.0);
    double Ad = (x + y + 3.0 * z) * 0.2;
    double sd = 0.0, f = 1.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        p[i] = p_in[i];
        Aj[i] = (x + y + z + 2.0 * p[i]) * 0.2;
        sj[i] = 0.0;
        dl[i] = (p[i] - x) * (p[i] - y) * (p[i] - z);   // delta_0; delta_n = delta_0 / 64^n exactly (Carlson 1995, eq. 2.19)
    }
    // Every deviation (mean - argument) shrinks by EXACTLY 4 per duplication step ((A+lam)/4 - (x+lam)/4 = (A-x)/4), so the
    // largest initial deviation Q0 is measured once and the test per step is Q0 * 4^-n < tol * (smallest mean).
    // a mean lies inside the range of its arguments, so its largest deviation is max(A - lowest, highest - A)
    double Q0 = 0.0, Amin = kInf;
    const double lo3 = fmin(fmin(x, y), z), hi3 = fmax(fmax(x, y), z);
    if (WANT_RF) { Q0 = fmax(Q0, fmax(Af - lo3, hi3 - Af)); Amin = fmin(Amin, Af); }
    if (WANT_RD) { Q0 = fmax(Q0, fmax(Ad - lo3, hi3 - Ad)); Amin = fmin(Amin, Ad); }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        Q0 = fmax(Q0, fmax(Aj[i] - fmin(lo3, p[i]), fmax(hi3, p[i]) - Aj[i]));
        Amin = fmin(Amin, Aj[i]);
    }
    if (!(Q0 == Q0) || !(Amin == Amin)) Q0 = 0.0;                     // NaN arguments: nothing to converge to
#pragma unroll 1
    for (int n = 0; n < kMaxDup; ++n) {
        if (Q0 * f < kBundleTol * Amin) break;
        double sx = fsqrt(x), sy = fsqrt(y), sz = fsqrt(z);
        double lam = fma(sx, sy, fma(sy, sz, sz * sx));
        if (WANT_RD) sd += fdiv(f, sz * (z + lam));
        const double f3 = f * f * f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            double sp = fsqrt(p[i]);
            double id = frcp((sp + sx) * (sp + sy) * (sp + sz));
            // e_n = (sp-sx)(sp-sy)(sp-sz)/d_n = delta_n / d_n^2
            double e = (dl[i] * f3) * (id * id);
            sj[i] += f * id * rc1p_lazy(e, sp, p[i] + lam, id);
            p[i] = (p[i] + lam) * 0.25;
            Aj[i] = (Aj[i] + lam) * 0.25;
        }
        if (WANT_RF) Af = (Af + lam) * 0.25;
        if (WANT_RD) Ad = (Ad + lam) * 0.25;
        Amin = (Amin + lam) * 0.25;
        x = (x + lam) * 0.25;
        y = (y + lam) * 0.25;
        z = (z + lam) * 0.25;
        f *= 0.25;
    }
    if (WANT_RF) {
        double ys = frsqrt(Af), iA = ys * ys;
        double X = (Af - x) * iA, Y = (Af - y) * iA, Z = -(X + Y);
        *rf = rf_series(X * Y - Z * Z, X * Y * Z) * ys;
    }
    if (WANT_RD) {
        double ys = frsqrt(Ad), iA = ys * ys;
        double X = (Ad - x) * iA, Y = (Ad - y) * iA, Z = -(X + Y) * (1.0 / 3.0);
        double XY = X * Y, Z2 = Z * Z;
        *rd = f * rj_series(XY - 6.0 * Z2, (3.0 * XY - 8.0 * Z2) * Z, 3.0 * (XY - Z2) * Z2, XY * Z2 * Z) * (iA * ys) + 3.0 * sd;
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double ys = frsqrt(Aj[i]), iA = ys * ys;
        double X = (Aj[i] - x) * iA, Y = (Aj[i] - y) * iA, Z = (Aj[i] - z) * iA, P = -(X + Y + Z) * 0.5;
        double XYZ = X * Y * Z, P2 = P * P;
        double E2 = X * Y + X * Z + Y * Z - 3.0 * P2;
        double E3 = XYZ + 2.0 * E2 * P + 4.0 * P2 * P;
        double E4 = (2.0 * XYZ + E2 * P + 3.0 * P2 * P) * P;
        double E5 = XYZ * P2;
        rj[i] = f * rj_series(E2, E3, E4, E5) * (iA * ys) + 6.0 * sj[i];
    }
}

// ---- Legendre pieces in "reduced" form ------------------------------------------------------------------------
// Inputs: s = sin(phi_r), c2 = cos^2(phi_r), y = 1 - m s^2 (|phi_r| <= pi/2), F = F(phi_r|m).
// pi_tail(n) = Pi(n;phi_r|m) - F(phi_r|m) = (n/3) s^3 RJ(c2, y, 1, 1 - n s^2), with the DLMF 19.20.14 principal value
// when rho = 1 - n s^2 < 0  (kgeo/scipy_ellip_binding.py:207-230).  rho == 0 -> +inf (:154).
struct PiPrep {
    double p;     // positive 4th argument handed to RJ
    double rho;   // 1 - n s^2
};
KG_HD PiPrep pi_prepare(double n, double s, double c2, double y) {
    PiPrep o;
    o.rho = fma(-n * s, s, 1.0);
    if (o.rho > 0.0) {
        o.p = o.rho;
    } else {
        double q = -o.rho;
        o.p = ((c2 + y + q) - c2 * y) / (1.0 + q);
    }
    return o;
}
// Cauchy principal value for rho < 0 (out of line: one copy for every Pi of the fused kernels)
KG_NOINLINE double pi_finish_pv(double s3n, double s, double c2, double y, double F, double p, double rho, double rj) {
    double q = -rho, xy = c2 * y, pq = p * q;
    double isq = frsqrt(xy + pq);                                  // xy + pq > 0 here
    double t = fsqrt(fmax(xy, 1.0e-300)) * isq;
    double ath = 0.5 * flog(fdiv(1.0 + t, 1.0 - t));                // = sqrt(xy) * RC(xy+pq, pq)
    double cpv = fdiv((p - 1.0) * rj - 3.0 * fdiv(F, s) + 3.0 * ath * isq, q + 1.0);
    return s3n * cpv;
}
KG_HD double pi_finish(double n, double s, double c2, double y, double F, const PiPrep& pp, double rj) {
    double s3n = (n * (1.0 / 3.0)) * cube(s);
    if (pp.rho > 0.0) return s3n * rj;
    if (pp.rho == 0.0) return kInf;
    return pi_finish_pv(s3n, s, c2, y, F, pp.p, pp.rho, rj);
}

// ---- complete integrals --------------------------------------------------------------------------------------
// complete integrals are setup-only: out of line, one copy per kernel (the setup code must stay inside the instruction cache)
KG_NOINLINE double ellip_K(double m) { return carlson_rf(0.0, 1.0 - m, 1.0); }
KG_NOINLINE double ellip_RDc(double m) { return carlson_rd(0.0, 1.0 - m, 1.0); }         // (K-E)/m = RDc/3
KG_HD double ellip_Ec(double m) { return ellip_K(m) - (m * (1.0 / 3.0)) * ellip_RDc(m); }
// Pi(n|m) - K(m); principal value for n > 1; n == 1 -> +inf
KG_NOINLINE double ellip_PicTail(double n, double m) {
    double y = 1.0 - m;
    if (n == 1.0) return kInf;
    PiPrep pp = pi_prepare(n, 1.0, 0.0, y);
    double rj = carlson_rj(0.0, y, 1.0, pp.p);
    if (pp.rho > 0.0) return (n * (1.0 / 3.0)) * rj;
    double q = -pp.rho;
    return (n * (1.0 / 3.0)) * ((pp.p - 1.0) * rj - 3.0 * ellip_K(m)) / (q + 1.0);
}

// K(m), RD(0, 1-m, 1) and NP tails Pi(n_i|m) - K(m) of ONE parameter m on ONE shared duplication track (x = 0, y = 1 - m, z = 1)
// with the 1-ulp seeded square roots / divisions of the per-sample code.  The per-ray setup needs 2 + NP complete integrals of
// each motion's parameter; evaluated one by one through ellip_K / ellip_RDc / ellip_PicTail (stand-alone Carlson routines, IEEE
// square roots and divisions, one duplication track EACH) they were three quarters of ray_setup_kernel's instructions.
// Same mathematics, results within an ulp or two of the stand-alone routines.  n_i == 1 gives +inf as ellip_PicTail does.
template <int NP, bool WANT_K = true, bool WANT_RD = true>
KG_NOINLINE void ellip_complete(double m, const double* n, double* K, double* RDc, double* PicT) {
    const double y = 1.0 - m;
    PiPrep pp[NP > 0 ? NP : 1];
    double pa[NP > 0 ? NP : 1], rj[NP > 0 ? NP : 1];
    bool need_K = WANT_K;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        pp[i] = pi_prepare(n[i], 1.0, 0.0, y);
        pa[i] = (n[i] == 1.0) ? 1.0 : pp[i].p;             // placeholder: the result is overwritten with +inf
    }
    double rf = 0.0, rd = 0.0;
    // the principal-value branch of a tail (rho < 0) needs K(m) as well: always taken along when there are tails
    carlson_bundle<NP, (WANT_K || NP > 0), WANT_RD>(0.0, y, pa, &rf, &rd, rj);
    (void)need_K;
    if (K) *K = rf;
    if (RDc) *RDc = rd;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double v;
        if (n[i] == 1.0) v = kInf;
        else if (pp[i].rho > 0.0) v = (n[i] * (1.0 / 3.0)) * rj[i];
        else v = (n[i] * (1.0 / 3.0)) * ((pp[i].p - 1.0) * rj[i] - 3.0 * rf) / (-pp[i].rho + 1.0);
        PicT[i] = v;
    }
}

// ---- incomplete Legendre integrals of a general amplitude (setup + element-wise API) ---------------------------
// Period reduction: phi = j*pi + phi_r, |phi_r| <= pi/2 (same mathematics as scipy_ellip_binding.py:164-192 and
// cephes ellik/ellie).  Returns F, E and up to NP Pi's for one (phi, m).
template <int NP>
KG_HD void legendre_all(double phi, double m, const double* n, double* F, double* E, double* Pi) {
    double j = rint(phi * (1.0 / kPi));
    double phir = fma(-j, kPi, phi);
    phir = fma(-j, 1.2246467991473532e-16, phir);   // low word of pi
    double s = sin(phir), c = cos(phir);
    double c2 = c * c, y = fma(-m * s, s, 1.0);
    PiPrep pp[NP > 0 ? NP : 1];
    double pa[NP > 0 ? NP : 1], rj[NP > 0 ? NP : 1];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        pp[i] = pi_prepare(n[i], s, c2, y);
        pa[i] = pp[i].p;
    }
    double rf, rd;
    carlson_bundle<NP, true, true>(c2, y, pa, &rf, &rd, rj);
    double Fr = s * rf;
    double Er = Fr - (m * (1.0 / 3.0)) * cube(s) * rd;
    double Kc = 0.0, RDc = 0.0;
    if (j != 0.0) {
        Kc = ellip_K(m);
        RDc = ellip_RDc(m);
    }
    *F = (j != 0.0) ? fma(2.0 * j, Kc, Fr) : Fr;
    *E = (j != 0.0) ? fma(2.0 * j, Kc - (m * (1.0 / 3.0)) * RDc, Er) : Er;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double tail = pi_finish(n[i], s, c2, y, Fr, pp[i], rj[i]);
        double v = Fr + tail;
        if (j != 0.0) v += 2.0 * j * (Kc + ellip_PicTail(n[i], m));
        Pi[i] = v;
    }
}

KG_HD double ellip_F(double phi, double m) {
    double F, E;
    legendre_all<0>(phi, m, nullptr, &F, &E, nullptr);
    return F;
}
KG_HD double ellip_E(double phi, double m) {
    double F, E;
    legendre_all<0>(phi, m, nullptr, &F, &E, nullptr);
    return E;
}

// Pi(n;phi|m) with the reference's special-value conventions, in the reference's precedence
// (kgeo/scipy_ellip_binding.py:125-162).
KG_HD double ellip_Pi_ref(double n, double phi, double m) {
    double aphi = fabs(phi);
    if (aphi <= 1.0e-10) return phi;
    if (m < -1.0e14) return 0.0;
    double s0 = sin(phi);
    if (1.0 - m * s0 * s0 < 0.0) return kg_nan();
    if (aphi > kHalfPi && m >= 1.0) return kg_nan();
    if (aphi > kHalfPi && n == 1.0) return kInf;
    if (1.0 - n * s0 * s0 == 0.0) return kInf;
    if (isinf(n)) return 0.0;
    double F, E, Pi;
    legendre_all<1>(phi, m, &n, &F, &E, &Pi);
    return Pi;
}

// ---- Jacobi elliptic functions, 0 <= m <= 1 ----------------------------------------------------------------------
// Descending Landen / AGM ladder (Bulirsch's sncndn).  Also returns K(m) = pi / (2 * agm(1, sqrt(1-m))).
constexpr int kAgmMax = 12;
KG_HD void jacobi_sncndn(double u, double m, double* sn, double* cn, double* dn, double* Kout) {
    double mc = 1.0 - m;
    if (!(mc > 0.0)) {  // m == 1 (or invalid): hyperbolic limit
        if (mc == 0.0) {
            double ch = cosh(u);
            *sn = tanh(u);
            *cn = 1.0 / ch;
            *dn = 1.0 / ch;
            *Kout = kInf;
        } else {
            *sn = *cn = *dn = *Kout = kg_nan();
        }
        return;
    }
    double la[kAgmMax], lg[kAgmMax];
    double a = 1.0, g, c = 1.0;
    int l = 0;
    for (int i = 0; i < kAgmMax; ++i) {
        l = i;
        la[i] = a;
        g = sqrt(mc);
        lg[i] = g;
        c = 0.5 * (a + g);
        if (fabs(a - g) <= 1.5e-8 * a) break;
        mc = g * a;
        a = c;
    }
    *Kout = kHalfPi / c;
    double uu = u * c;
    double s, co;
#if defined(__CUDA_ARCH__)
    sincos(uu, &s, &co);
#else
    s = sin(uu);
    co = cos(uu);
#endif
    double d = 1.0;
    if (s != 0.0) {
        double aa = co / s;
        double cc = c * aa;
        for (int i = l; i >= 0; --i) {
            double b = la[i];
            aa *= cc;
            cc *= d;
            d = (lg[i] + aa) / (b + aa);
            aa = cc / b;
        }
        aa = 1.0 / sqrt(fma(cc, cc, 1.0));
        s = (s >= 0.0) ? aa : -aa;
        co = cc * s;
    }
    *sn = s;
    *cn = co;
    *dn = d;
}

// ---- Jacobi functions from a precomputed AGM ladder ---------------------------------------------------------------
// The ladder (a_i, g_i) depends only on the parameter m, i.e. on the RAY, not on the Mino time: the per-ray setup
// stores it once and every (ray, sample) evaluation skips the ~6 square roots of the descent.
// Layout: L[0] = index of the last level (or -1 for m == 1), L[1] = c = last arithmetic mean,
//         L[2+3i] = a_i, L[3+3i] = 1/a_i, L[4+3i] = g_i.
// Pointer into a slot-major workspace: slot s of this ray lives at base[s * stride].  The per-ray setup is written against
// a generic pointer type P (double* for a thread-local array, WsPtr for the [slot][ray] global workspace), so the setup
// kernel stores every constant straight into HBM instead of staging ~2 KB per thread in local memory.
struct WsPtr {
    double* base;
    size_t stride;
    KG_HD double& operator[](int s) const { return base[(size_t)s * stride]; }
    KG_HD WsPtr operator+(int s) const { return WsPtr{base + (size_t)s * stride, stride}; }
};

constexpr int kLadderLevels = 10;
constexpr int kLadderSize = 2 + 3 * kLadderLevels;

template <class P>
KG_NOINLINE void jacobi_make_ladder(double m, P L) {
    for (int i = 0; i < kLadderSize; ++i) L[i] = 0.0;
    double mc = 1.0 - m;
    if (!(mc > 0.0)) {
        L[0] = -1.0;
        L[1] = (mc == 0.0) ? 1.0 : kg_nan();
        return;
    }
    double a = 1.0, g, c = 1.0;
    int l = 0;
    for (int i = 0; i < kLadderLevels; ++i) {
        l = i;
        g = sqrt(mc);
        L[2 + 3 * i] = a;
        L[3 + 3 * i] = 1.0 / a;
        L[4 + 3 * i] = g;
        c = 0.5 * (a + g);
        if (fabs(a - g) <= 1.5e-8 * a) break;
        mc = g * a;
        a = c;
    }
    L[0] = (double)l;
    L[1] = c;
}

// out-of-line libm calls shared by every call site of the fused kernels (code size: the track kernel must stay
// inside the instruction cache)
KG_NOINLINE double kg_log(double x) { return flog(x); }
KG_NOINLINE double kg_acos(double x) { return facos(x); }

struct SnCnDn {
    double sn, cn, dn;
};

KG_NOINLINE void jacobi_m1(double u, double* sn, double* cn, double* dn) {
    double ch = cosh(u);
    *sn = tanh(u);
    *cn = 1.0 / ch;
    *dn = 1.0 / ch;
}

template <class A>
KG_NOINLINE SnCnDn jacobi_from_ladder(const A L, double u) {
    SnCnDn o;
    const int l = (int)L(0);
    const double c = L(1);
    if (l < 0) {   // m == 1: hyperbolic limit
        jacobi_m1(u, &o.sn, &o.cn, &o.dn);
        return o;
    }
    double uu = u * c;
    double s, co;
    fsincos(uu, &s, &co);
    double d = 1.0;
    if (s != 0.0) {
        double aa = fdiv(co, s);
        double cc = c * aa;
#pragma unroll 1
        for (int i = l; i >= 0; --i) {
            const double b = L(2 + 3 * i), ib = L(3 + 3 * i), g = L(4 + 3 * i);
            aa *= cc;
            cc *= d;
            d = fdiv(g + aa, b + aa);
            aa = cc * ib;
        }
        aa = frsqrt(fma(cc, cc, 1.0));
        s = (s >= 0.0) ? aa : -aa;
        co = cc * s;
    }
    o.sn = s;
    o.cn = co;
    o.dn = d;
    return o;
}

// amplitude am(u|m), continuous in u (scipy.special.ellipj's 4th output)
KG_HD double jacobi_am(double u, double sn, double cn, double K) {
    if (!(K < kInf)) return atan2(sn, cn);   // m == 1: am = gd(u) in (-pi/2, pi/2)
    double n = rint(u / (4.0 * K));
    double ur = fma(-4.0 * K, n, u);
    double am = atan2(sn, cn);                 // in (-pi, pi]
    if (ur > 0.0 && am < -kHalfPi) am += 2.0 * kPi;
    if (ur < 0.0 && am > kHalfPi) am -= 2.0 * kPi;
    return fma(2.0 * kPi, n, am);
}

}  // namespace kg
