// Equatorial-image physics: crossing bookkeeping, photon momentum signs, redshift, fluid 4-velocity models, magnetic
// field models, polarization (tetrads + Penrose-Walker), emissivities.  Everything is evaluated in the equatorial plane.
//
// Replaces kgeo/equatorial_lensing.py:85-196 (r_equatorial), kgeo/equatorial_images.py:124-544 (Iobs and helpers),
// kgeo/velocities.py:70-470, kgeo/bfields.py:93-172,330-396,429-542,712-746, kgeo/emissivities.py:83-157.
#pragma once
#include "kg_geodesic.cuh"
#include "../../include/kgeo_b200.h"

namespace kg {

// The reference evaluates cos(pi/2) in floating point (6.1e-17) instead of 0; it only matters where it enters
// linearly (Bfield_simple's vertical term), so it is kept there and dropped where it is squared.
constexpr double kCosHalfPi = 6.123233995736766e-17;

struct U4 {
    double u0, u1, u2, u3;
};
struct EqMetric {
    double Delta, Sigma, g00, g11, g22, g33, g03;
};

KG_HD EqMetric eq_metric(double a, double r) {
    // kgeo/velocities.py:102-114 at theta = pi/2
    EqMetric g;
    double a2 = a * a, r2 = r * r;
    g.Delta = r2 - 2.0 * r + a2;
    g.Sigma = r2;
    g.g00 = -(1.0 - 2.0 * r / g.Sigma);
    g.g11 = g.Sigma / g.Delta;
    g.g22 = g.Sigma;
    g.g33 = (r2 + a2 + 2.0 * r * a2 / g.Sigma);
    g.g03 = -2.0 * r * a / g.Sigma;
    return g;
}

KG_HD double isco_radius(double a, double s) {
    double z1 = 1.0 + cbrt(1.0 - a * a) * (cbrt(1.0 + a) + cbrt(1.0 - a));
    double z2 = sqrt(3.0 * a * a + z1 * z1);
    return 3.0 + z2 - s * sqrt((3.0 - z1) * (3.0 + z1 + 2.0 * z2));
}
KG_HD double sgn0(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }   // np.sign
KG_HD double pow15(double x) { return x * sqrt(x); }

KG_HD U4 u_zamo(double a, double r) {   // velocities.py:179-208
    EqMetric g = eq_metric(a, r);
    double v3 = -g.g03 / g.g33;
    double u0 = sqrt(-1.0 / (g.g00 + 2.0 * g.g03 * v3 + g.g33 * v3 * v3));
    return U4{u0, 0.0, 0.0, u0 * v3};
}
KG_HD U4 u_infall(double a, double r) {   // velocities.py:211-226
    double r2 = r * r, a2 = a * a;
    double Delta = r2 + a2 - 2.0 * r;
    double Xi = sq(r2 + a2) - Delta * a2;
    return U4{Xi / (r2 * Delta), -sqrt(2.0 * r * (r2 + a2)) / r2, 0.0, 2.0 * a / (r * Delta)};
}
KG_HD U4 u_kep(double a, double r, bool retro) {   // velocities.py:229-281
    double s = retro ? -1.0 : 1.0;
    double ri = isco_radius(a, s);
    double spin = fabs(a), asign = sgn0(a);
    U4 u{0.0, 0.0, 0.0, 0.0};
    if (r >= ri) {
        double r15 = pow15(r);
        u.u0 = (r15 + s * spin) / sqrt(r * r * r - 3.0 * r * r + 2.0 * s * spin * r15);
        u.u3 = (asign * s / (r15 + s * spin)) * u.u0;
    } else {
        double Delta = r * r - 2.0 * r + a * a;
        double ell_i = s * asign * (ri * ri + a * a - s * 2.0 * spin * sqrt(ri)) / (pow15(ri) - 2.0 * sqrt(ri) + s * spin);
        double gam_i = sqrt(1.0 - 2.0 / (3.0 * ri));
        double H = (2.0 * r - a * ell_i) / Delta;
        u.u0 = gam_i * (1.0 + (2.0 / r) * (1.0 + H));
        u.u1 = -sqrt(2.0 / (3.0 * ri)) * pow15(ri / r - 1.0);
        u.u3 = gam_i * (ell_i + a * H) / (r * r);
    }
    return u;
}
KG_HD U4 u_subkep(double a, double r, bool retro, double fac) {   // velocities.py:284-359
    double s = retro ? -1.0 : 1.0;
    double ri = isco_radius(a, s);
    double spin = fabs(a), asign = sgn0(a);
    double a2 = a * a;
    U4 u{0.0, 0.0, 0.0, 0.0};
    double Delta = r * r - 2.0 * r + a2;
    double Xi = sq(r * r + a2) - Delta * a2;
    if (r >= ri) {
        double ell = asign * s * (r * r + spin * spin - s * 2.0 * spin * sqrt(r)) / (pow15(r) - 2.0 * sqrt(r) + s * spin);
        ell *= fac;
        double gam = sqrt(Delta / (Xi / (r * r) - 4.0 * a * ell / r - (1.0 - 2.0 / r) * ell * ell));
        double H = (2.0 * r - a * ell) / Delta;
        double chi = 1.0 / (1.0 + (2.0 / r) * (1.0 + H));
        double Omega = (chi / (r * r)) * (ell + a * H);
        u.u0 = gam / chi;
        u.u3 = (gam / chi) * Omega;
    } else {
        double Delta_i = ri * ri - 2.0 * ri + a2;
        double Xi_i = sq(ri * ri + a2) - Delta_i * a2;
        double ell_i = asign * s * (ri * ri + spin * spin - s * 2.0 * spin * sqrt(ri)) / (pow15(ri) - 2.0 * sqrt(ri) + s * spin);
        ell_i *= fac;
        double gam_i = sqrt(Delta_i / (Xi_i / (ri * ri) - 4.0 * a * ell_i / ri - (1.0 - 2.0 / ri) * ell_i * ell_i));
        double H = (2.0 * r - a * ell_i) / Delta;
        double chi = 1.0 / (1.0 + (2.0 / r) * (1.0 + H));
        double Omega = (chi / (r * r)) * (ell_i + a * H);
        double nu = (r / Delta) * sqrt(Xi / (r * r) - 4.0 * a * ell_i / r - (1.0 - 2.0 / r) * ell_i * ell_i - Delta / (gam_i * gam_i));
        u.u0 = gam_i / chi;
        u.u1 = -gam_i * (Delta / (r * r)) * nu;
        u.u3 = (gam_i / chi) * Omega;
    }
    return u;
}
KG_HD U4 u_gelles(double a, double r, double beta, double chi) {   // velocities.py:362-389
    double a2 = a * a, r2 = r * r;
    double Delta = r2 - 2.0 * r + a2;
    double Xi = sq(r2 + a2) - Delta * a2;
    double omegaz = 2.0 * a * r / Xi;
    double gamma = 1.0 / sqrt(1.0 - beta * beta);
    U4 u;
    u.u0 = (gamma / r) * sqrt(Xi / Delta);
    u.u1 = (beta * gamma * cos(chi) / r) * sqrt(Delta);
    u.u2 = 0.0;
    u.u3 = (gamma * omegaz / r) * sqrt(Xi / Delta) + (r * beta * gamma * sin(chi)) / sqrt(Xi);
    return u;
}
KG_HD U4 u_simfit(double a, double r, double ell_isco, double vr_isco, double p1, double p2, double dd) {   // :392-439
    double r_isco = isco_radius(a, 1.0);
    double a2 = a * a, r2 = r * r;
    double Delta = r2 - 2.0 * r + a2, Sigma = r2;
    double g00u = -(r2 + a2 + 2.0 * r * a2 / Sigma) / Delta;
    double g11u = Delta / Sigma;
    double g33u = (Delta - a2) / (Sigma * Delta);
    double g03u = -(2.0 * r * a) / (Sigma * Delta);
    double x = r / r_isco;
    double ell = ell_isco * sqrt(x);
    double vr = -vr_isco * pow(x, -p1) * pow(0.5 * (1.0 + pow(x, 1.0 / dd)), (p1 - p2) * dd);
    double gam = sqrt(-1.0 / (g00u + g11u * vr * vr + g33u * ell * ell - 2.0 * g03u * ell));
    double u_0 = -gam, u_1 = vr * gam, u_3 = ell * gam;
    return U4{g00u * u_0 + g03u * u_3, g11u * u_1, 0.0, g33u * u_3 + g03u * u_0};
}
KG_HD U4 u_general(double a, double r, bool retro, double fac, double beta_phi, double beta_r) {   // :442-470
    double Delta = r * r + a * a - 2.0 * r;
    U4 ui = u_infall(a, r);
    U4 us = u_subkep(a, r, retro, fac);
    double Oi = ui.u3 / ui.u0, Os = us.u3 / us.u0;
    double u1 = us.u1 + (1.0 - beta_r) * (ui.u1 - us.u1);
    double Omega = Os + (1.0 - beta_phi) * (Oi - Os);
    double u0 = sqrt(1.0 + (r * r) * (u1 * u1) / Delta);
    u0 /= sqrt(1.0 - (r * r + a * a) * Omega * Omega - (2.0 / r) * sq(1.0 - a * Omega));
    return U4{u0, u1, 0.0, u0 * Omega};
}
// numpy.interp(x, xp, fp) for one x: linear between the bracketing samples, clamped to fp[0] / fp[n-1] outside, NaN -> NaN
// (numpy/core/src/multiarray/compiled_base.c: slope * (x - xp[j]) + fp[j])
KG_HD double interp_table(const double* xp, const double* fp, long long n, double x) {
    if (!(x == x)) return x;
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    long long lo = 0, hi = n - 1;          // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        long long mid = (lo + hi) >> 1;
        if (x >= xp[mid]) lo = mid;
        else hi = mid;
    }
    const double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
    return slope * (x - xp[lo]) + fp[lo];
}

// four-velocity from tabulated primitives U1, U2, U3 (kgeo/velocities.py:659-698), equatorial plane
KG_HD U4 u_fromfile(const kgeo_models& M, double a, double r) {
    const double* tab = M.velocity_table_dev;
    const long long n = M.velocity_table_n;
    double u1 = interp_table(tab, tab + n, n, r);
    double u2 = interp_table(tab, tab + 2 * n, n, r);
    double u3 = interp_table(tab, tab + 3 * n, n, r);
    double r2 = r * r, a2 = a * a;
    double Delta = r2 - 2.0 * r + a2, Sigma = r2;
    double g11 = Sigma / Delta, g22 = Sigma, g33 = r2 + a2 + 2.0 * r * a2 / Sigma;
    double gcon00 = -(r2 + a2 + 2.0 * r * a2 / Sigma) / Delta;
    double gcon03 = -2.0 * r * a / Sigma / Delta;
    double alpha = 1.0 / sqrt(-gcon00);
    double gamma = sqrt(1.0 + g11 * u1 * u1 + g22 * u2 * u2 + g33 * u3 * u3);
    return U4{gamma / alpha, u1, u2, u3 - gamma * alpha * gcon03};
}

// Principal branch of the Lambert W function for real x >= -1/e (scipy.special.lambertw(x, k=0) on that range): series
// about the branch point or a logarithmic / rational start, then Fritsch's iteration (quartic convergence, stable next to
// W = -1 because it works with ln(x/w) - w instead of w e^w - x)
KG_HD double lambertw0(double x) {
    const double em1 = 0.36787944117144233;   // 1/e
    if (!(x >= -em1)) return kg_nan();
    if (x == 0.0) return 0.0;
    double w;
    const double pp = 2.0 * (2.718281828459045 * x + 1.0);
    if (x < -0.25) {
        const double q = sqrt(pp > 0.0 ? pp : 0.0);
        w = -1.0 + q * (1.0 + q * (-1.0 / 3.0 + q * (11.0 / 72.0 + q * (-43.0 / 540.0 + q * (769.0 / 17280.0 + q * (-221.0 / 8505.0))))));
        if (q < 1.0e-6) return w;
    } else if (x < 3.0) {
        w = x * (1.0 + (4.0 / 3.0) * x) / (1.0 + x * (7.0 / 3.0 + (5.0 / 6.0) * x));   // Pade start about 0
    } else {
        const double l1 = log(x), l2 = log(l1);
        w = l1 - l2 + l2 / l1;
    }
    for (int it = 0; it < 4; ++it) {
        const double z = log(x / w) - w;
        const double w1 = 1.0 + w;
        const double q = 2.0 * w1 * (w1 + (2.0 / 3.0) * z);
        const double e = (z / w1) * ((q - z) / (q - 2.0 * z));
        w = w * (1.0 + e);
        if (fabs(e) < 1.0e-16) break;
    }
    return w;
}

// piecewise-linear table f(x) with FITPACK's k=1, ext=0 semantics (scipy UnivariateSpline(k=1, s=0): linear extrapolation
// with the end segments) and its derivative (slope of the segment holding x; right-continuous at the knots)
KG_HD void linear_spline(const double* xp, const double* fp, long long n, double x, double* f, double* fprime) {
    long long lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        long long mid = (lo + hi) >> 1;
        if (x >= xp[mid]) lo = mid; else hi = mid;
    }
    const double h = xp[lo + 1] - xp[lo];
    *f = (fp[lo] * (xp[lo + 1] - x) + fp[lo + 1] * (x - xp[lo])) / h;
    *fprime = (fp[lo + 1] - fp[lo]) / h;
}

// lab-frame B^i at theta = pi/2 and the field-line angular speed (kgeo/bfields.py:93-131,175-197 and the model functions);
// omega = NaN for the models the reference's omega_field does not cover
KG_HD void bfield_lab_omega(const kgeo_models& M, double a, double r, double* B1, double* B2, double* B3, double* omega) {
    const double* p = M.bfield_p;
    double a2 = a * a, r2 = r * r;
    double Delta = r2 - 2.0 * r + a2;
    *omega = kg_nan();
    switch (M.bfield_type) {
        case KGEO_B_BZ_MONOPOLE: {   // bfields.py:545-592; cos(pi/2) = 6.1e-17 kept as in the reference
            const double C = p[0];
            const bool second = p[1] != 0.0;
            const double* tab = M.bfield_table_dev;
            const long long n = M.bfield_table_n;
            double fr, frp;
            linear_spline(tab, tab + n, n, r, &fr, &frp);
            if (r > tab[n - 1]) {   // RMAXINTERP
                fr = 1.0 / (4.0 * r);
                frp = -1.0 / (4.0 * r2);
            }
            const double cth = kCosHalfPi, cth2 = cth * cth;
            const double gdet = r2 + a2 * cth2;
            const double cos2th = -1.0;
            const double dphidtheta = C * (1.0 + 0.5 * a2 * fr * (1.0 + 3.0 * cos2th));
            const double dphidr = C * a2 * cth * frp;
            const double f2 = (6.0 * kPi * kPi - 49.0) / 72.0;
            const double omega2 = 1.0 / 32.0 - (4.0 * f2 - 1.0) / 64.0;
            const double I = -1.0 * C * 2.0 * kPi * (a / 8.0 + (a * a * a) * (omega2 + 0.25 * fr * cth2));
            double Br = dphidtheta / gdet, Bth = -dphidr / gdet, Bph = I / (2.0 * kPi * Delta);
            if (second) {
                const double rr = r2 - 2.0 * r;
                Br = C / r2 + C * a2 * (-cth2 / (r2 * r2) + fr * (0.5 + 1.5 * cos2th) / r2);
                Bth = -C * a2 * frp * cth / r2;
                Bph = -C * 0.125 * a / rr + C * (a * a * a) * (0.125 / (rr * rr) - (omega2 + 0.25 * cth2 * fr) / rr);
            }
            const double spin = fabs(a);
            double Om = spin / 8.0 + (spin * spin * spin) * omega2;
            if (a < 0.0) Om = -Om;
            *B1 = Br; *B2 = Bth; *B3 = Bph; *omega = Om;
            return;
        }
        case KGEO_B_BZ_PARA: {   // bfields.py:609-673 (array branch: omega carries the factor C and the argvals mask)
            const double C = p[0];
            const double rp = 1.0 + sqrt(1.0 - a2);
            const double c = kCosHalfPi;
            const double gdet = r2 + a2 * c * c;
            const double psi = r * (1.0 - c) + rp * (1.0 + c) * (1.0 - log(1.0 + c)) - 2.0 * rp * (1.0 - 0.6931471805599453);
            const double dpsidtheta = r + rp * log(1.0 + c);
            const double dpsidr = 1.0 - c;
            const double arg = -psi / rp + 1.3862943611198906;   // + log(4)
            double Om = 0.0;
            if (arg >= -0.36787944117144233) {
                const double wf = lambertw0(arg);
                const double xf = exp(wf);
                const double yf = 1.0 + psi / (1.0 + wf) / (xf * (2.0 - xf));
                Om = C * (a / 4.0 / yf);
            }
            const double I = -4.0 * kPi * psi * Om;
            *B1 = C * dpsidtheta / gdet;
            *B2 = -C * dpsidr / gdet;
            *B3 = I / (2.0 * kPi * Delta);
            *omega = Om;
            return;
        }
        case KGEO_B_POWER: {   // bfields.py:618-625,676-709
            const double pv = p[0], C = p[1];
            const bool usemono = p[2] != 0.0;
            const double rp = 1.0 + sqrt(1.0 - a2);
            const double c = kCosHalfPi;
            const double gdet = r2 + a2 * c * c;
            const double rpw = pow(r, pv);
            const double psi = rpw * (1.0 - c);
            const double dpsidtheta = rpw;
            const double dpsidr = pv * psi / r;
            double Om;
            if (usemono || pv == 0.0) {
                Om = a / 8.0;
            } else {
                const double cthh = 1.0 - psi * pow(rp, -pv);
                Om = a / (4.0 + 8.0 / (1.0 + cthh));
            }
            const double I = (pv > 0.0) ? (-4.0 * kPi * psi * Om) : (-2.0 * kPi * psi * (2.0 - psi) * Om);
            *B1 = C * dpsidtheta / gdet;
            *B2 = -C * dpsidr / gdet;
            *B3 = I / (2.0 * kPi * Delta);
            *omega = Om;
            return;
        }
        case KGEO_B_SIMPLE:
        case KGEO_B_SIMPLE_RM1: {   // bfields.py:429-477
            double gdet = r2;
            double rad = (M.bfield_type == KGEO_B_SIMPLE_RM1) ? (r / gdet) : (1.0 / gdet);
            *B1 = p[0] * rad + p[1] * (r * kCosHalfPi / gdet);
            *B2 = p[1] * (-1.0 / gdet);
            *B3 = p[2] * (1.0 / gdet);
            return;
        }
        case KGEO_B_BZ_GUESS: {   // bfields.py:504-542
            double Pi = sq(r2 + a2) - a2 * Delta;
            double rH = 1.0 + sqrt(1.0 - a2);
            double OmegaH = a / (2.0 * rH);
            double spin = fabs(a);
            double f2 = (6.0 * kPi * kPi - 49.0) / 72.0;
            double omega2 = 1.0 / 32.0 - (4.0 * f2 - 1.0) / 64.0;
            double OmegaBZ = spin / 8.0 + (spin * spin * spin) * omega2;
            if (a < 0.0) OmegaBZ = -OmegaBZ;
            double ratio = -1.0 * (OmegaH - OmegaBZ) * sqrt(Pi) / Delta;
            double Br = p[0] / r2;
            *B1 = Br;
            *B2 = 0.0 * Br;
            *B3 = ratio * Br;
            *omega = OmegaBZ;
            return;
        }
        case KGEO_B_GEN_POWER: {   // bfields.py:712-746
            double n_I = p[0], p_val = p[1];
            double sg = (p[2] != 0.0) ? 1.0 : -1.0;
            double rp = 1.0 + sqrt(1.0 - a2);
            double gdet = r2;
            double I_0 = sg * (a * kPi) / (rp * rp);
            double dpsidtheta = sg * pow(r / rp, p_val);
            double I = I_0 * pow(r / rp, -1.0 * n_I);
            double dpsidr = (1.0 - kCosHalfPi) * ((p_val * pow(r, p_val - 1.0)) / pow(rp, p_val));
            *B1 = dpsidtheta / gdet;
            *B2 = -1.0 * (dpsidr / gdet);
            *B3 = I / (2.0 * kPi * Delta);
            *omega = 0.5 * (a / (2.0 * rp));
            return;
        }
        case KGEO_B_FROMFILE: {   // bfields.py:790-806
            const double* tab = M.bfield_table_dev;
            const long long n = M.bfield_table_n;
            *B1 = interp_table(tab, tab + n, n, r);
            *B2 = interp_table(tab, tab + 2 * n, n, r);
            *B3 = interp_table(tab, tab + 3 * n, n, r);
            return;
        }
        default: {   // KGEO_B_MONOPOLE_A, bfields.py:480-501
            double C = p[0], omegafac = p[1];
            double rh = 1.0 + sqrt(1.0 - a2);
            double omegah = a / (2.0 * rh);
            double phih = 4.0 * kPi * C;
            double omegaf = omegafac * omegah;
            double I = phih * rh * (omegaf - omegah) / (rh * rh);
            *B1 = (phih / (4.0 * kPi)) / r2;
            *B2 = 0.0;
            *B3 = I / (2.0 * kPi * Delta);
            *omega = omegaf;
            return;
        }
    }
}
KG_HD void bfield_lab(const kgeo_models& M, double a, double r, double* B1, double* B2, double* B3) {
    double omega;
    bfield_lab_omega(M, a, r, B1, B2, B3, &omega);
}

// drift-frame velocity of the image's B-field model (kgeo/velocities.py:487-623 with a numeric nu_parallel and no gammamax)
KG_HD U4 u_driftframe_B(double a, double r, double nu_parallel, double B1, double B2, double B3, double omegaf) {
    const double a2 = a * a, r2 = r * r;
    const double cth = kCosHalfPi, cth2 = cth * cth;
    const double Delta = r2 - 2.0 * r + a2;
    const double Sigma = r2 + a2 * cth2;
    const double Xi = sq(r2 + a2) - Delta * a2;
    const double omegaz = 2.0 * a * r / Xi;
    const double gdet = Sigma;
    const double g11 = Sigma / Delta, g22 = Sigma, g33 = Xi / Sigma;
    const double alpha = sqrt(Delta * Sigma / Xi);
    const double eta3 = 2.0 * a * r / sqrt(Delta * Sigma * Xi);
    const double E1 = (omegaf - omegaz) * Xi * B2 / Sigma;
    const double E2 = -(omegaf - omegaz) * Xi * B1 / (Sigma * Delta);
    const double E3 = 0.0;
    const double Bsq = g11 * B1 * B1 + g22 * B2 * B2 + g33 * B3 * B3;
    const double Esq = g11 * E1 * E1 + g22 * E2 * E2 + g33 * E3 * E3;
    const double B1c = g11 * B1, B2c = g22 * B2, B3c = g33 * B3;
    const double E1c = g11 * E1, E2c = g22 * E2, E3c = g33 * E3;
    const double pre = alpha / (Bsq * gdet);
    const double vperp1 = pre * (E2c * B3c - B2c * E3c);
    const double vperp2 = pre * (E3c * B1c - B3c * E1c);
    const double vperp3 = pre * (E1c * B2c - B1c * E2c);
    const double vpar_max = sqrt(1.0 - Esq / Bsq);
    const double sB = sqrt(Bsq);
    const double v1 = vperp1 + nu_parallel * vpar_max * B1 / sB;
    const double v2 = vperp2 + nu_parallel * vpar_max * B2 / sB;
    const double v3 = vperp3 + nu_parallel * vpar_max * B3 / sB;
    const double vsq = g11 * v1 * v1 + g22 * v2 * v2 + g33 * v3 * v3;
    const double gamma = 1.0 / sqrt(1.0 - vsq);
    return U4{gamma / alpha, gamma * v1, gamma * v2, gamma * (v3 + eta3)};
}

KG_HD U4 u_driftframe(const kgeo_models& M, double a, double r, double nu_parallel) {
    double B1, B2, B3, omegaf;
    bfield_lab_omega(M, a, r, &B1, &B2, &B3, &omegaf);
    return u_driftframe_B(a, r, nu_parallel, B1, B2, B3, omegaf);
}

// the closed-form velocity models (everything but the drift frame, which needs the B-field model)
KG_HD U4 velocity_closed_form(const kgeo_models& M, double a, double r) {
    const double* p = M.velocity_p;
    switch (M.velocity_type) {
        case KGEO_VEL_ZAMO: return u_zamo(a, r);
        case KGEO_VEL_INFALL: return u_infall(a, r);
        case KGEO_VEL_KEP: return u_kep(a, r, p[0] != 0.0);
        case KGEO_VEL_SUBKEP: return u_subkep(a, r, p[0] != 0.0, p[1]);
        case KGEO_VEL_GENERAL: return u_general(a, r, p[0] != 0.0, p[1], p[2], p[3]);
        case KGEO_VEL_GELLES: return u_gelles(a, r, p[0], p[1]);
        case KGEO_VEL_FROMFILE: return u_fromfile(M, a, r);
        default: return u_simfit(a, r, p[0], p[1], p[2], p[3], p[4]);
    }
}

// Velocity and lab-frame field of the models at one equatorial radius.  The image kernel walks this ONE copy of the model code
// (once per crossing, twice for the thermal emissivity's normalisation radius) instead of inlining the model switches at each
// of their five call sites: the kernel was 34 k instructions, and instruction fetch its top stall.
struct FluidAt {
    U4 u;
    double B1, B2, B3, omega;
};
KG_HD FluidAt fluid_at(const kgeo_models& M, double a, double r, bool need_B) {
    FluidAt f;
    f.B1 = f.B2 = f.B3 = 0.0;
    f.omega = kg_nan();
    const bool drift = (M.velocity_type == KGEO_VEL_DRIFTFRAME);
    if (need_B || drift) bfield_lab_omega(M, a, r, &f.B1, &f.B2, &f.B3, &f.omega);
    f.u = drift ? u_driftframe_B(a, r, M.velocity_p[0], f.B1, f.B2, f.B3, f.omega) : velocity_closed_form(M, a, r);
    return f;
}
// b^2 in the fluid frame from B^i and u^mu at the same radius (kgeo/bfields.py:330-396)
KG_HD double bsq_of(const EqMetric& g, const FluidAt& f) {
    const U4& u = f.u;
    double l1 = g.g11 * u.u1, l2 = g.g22 * u.u2, l3 = g.g33 * u.u3 + g.g03 * u.u0;
    double b0 = f.B1 * l1 + f.B2 * l2 + f.B3 * l3;
    double b1 = (f.B1 + b0 * u.u1) / u.u0, b2 = (f.B2 + b0 * u.u2) / u.u0, b3 = (f.B3 + b0 * u.u3) / u.u0;
    return b0 * (g.g00 * b0 + g.g03 * b3) + b1 * (g.g11 * b1) + b2 * (g.g22 * b2) + b3 * (g.g33 * b3 + g.g03 * b0);
}

KG_HD U4 velocity_u_lab(const kgeo_models& M, double a, double r) {   // velocities.py:70-96
    const double* p = M.velocity_p;
    switch (M.velocity_type) {
        case KGEO_VEL_ZAMO: return u_zamo(a, r);
        case KGEO_VEL_INFALL: return u_infall(a, r);
        case KGEO_VEL_KEP: return u_kep(a, r, p[0] != 0.0);
        case KGEO_VEL_SUBKEP: return u_subkep(a, r, p[0] != 0.0, p[1]);
        case KGEO_VEL_GENERAL: return u_general(a, r, p[0] != 0.0, p[1], p[2], p[3]);
        case KGEO_VEL_GELLES: return u_gelles(a, r, p[0], p[1]);
        case KGEO_VEL_FROMFILE: return u_fromfile(M, a, r);
        case KGEO_VEL_DRIFTFRAME: return u_driftframe(M, a, r, p[0]);
        default: return u_simfit(a, r, p[0], p[1], p[2], p[3], p[4]);
    }
}


// b^2 in the fluid frame, ideal MHD (kgeo/bfields.py:330-396)
KG_HD double bfield_bsq(const kgeo_models& M, double a, double r, const U4& u) {
    EqMetric g = eq_metric(a, r);
    double B1, B2, B3;
    bfield_lab(M, a, r, &B1, &B2, &B3);
    double l1 = g.g11 * u.u1, l2 = g.g22 * u.u2, l3 = g.g33 * u.u3 + g.g03 * u.u0;
    double b0 = B1 * l1 + B2 * l2 + B3 * l3;
    double b1 = (B1 + b0 * u.u1) / u.u0, b2 = (B2 + b0 * u.u2) / u.u0, b3 = (B3 + b0 * u.u3) / u.u0;
    return b0 * (g.g00 * b0 + g.g03 * b3) + b1 * (g.g11 * b1) + b2 * (g.g22 * b2) + b3 * (g.g33 * b3 + g.g03 * b0);
}

// x ** p as NumPy evaluates it for the exponents an image actually uses: NumPy's power loop squares for p == 2 and libm's pow is
// exact for p == 1, 2 (the default spectral index 1 gives g ** 1 and sin(theta_b) ** 2); everything else goes to pow.  Two pow
// calls per order were ~13 % of the unpolarized image kernel's instructions.
KG_HD double pow_small_int(double x, double p) {
    if (p == 1.0) return x;
    if (p == 2.0) return x * x;
    return pow(x, p);
}

// rest-frame emissivity (kgeo/emissivities.py:83-157); has_g / has_s mirror the `is not None` tests
KG_HD double emissivity_jrest(const kgeo_models& M, double a, double r, bool has_g, double g, bool has_s, double sinthb,
                              double Bmag, double nu_obs, double specind) {
    const double* p = M.emissivity_p;
    double j;
    switch (M.emissivity_type) {
        case KGEO_EM_THERMAL: {
            const double e = 4.80320425e-10, me = 9.10938356e-28, c = 2.99792458e10, kB = 1.380649e-16;
            double nu_em = nu_obs / g;
            double x = r / p[0];
            double ne = p[1] * pow(x, -p[4]);
            double Te = p[2] * pow(x, -p[5]);
            double B = (p[7] != 0.0) ? Bmag : p[3] * pow(x, -p[6]);
            double Theta_e = kB * Te / (me * c * c);
            double nu_c = 1.5 * (e * B / (2.0 * kPi * me * c)) * (Theta_e * Theta_e) * sinthb;
            double xx = nu_em / fmax(nu_c, 1e-40);
            double x13 = cbrt(xx);
            double II = 2.5651 * (1.0 + 1.92 / x13 + 0.9977 / (x13 * x13)) * exp(-1.8899 * x13);
            return (ne * (e * e) * nu_em) / (2.0 * sqrt(3.0) * c * (Theta_e * Theta_e)) * II;
        }
        case KGEO_EM_CONSTANT: j = 1.0; break;
        case KGEO_EM_BPL: {
            double rh = 1.0 + sqrt(1.0 - a * a);
            double lg = log(r / rh);
            j = exp(p[0] * lg + p[1] * (lg * lg));
            break;
        }
        default: {   // GLM / RING
            double mu, goff, sig;
            if (M.emissivity_type == KGEO_EM_RING) {
                mu = p[0]; goff = p[1]; sig = p[2];
            } else {
                mu = 0.0; goff = p[0]; sig = p[1];
            }
            if (mu == 0.0) mu = 1.0 - sqrt(1.0 - a * a);   // `if mu_ring:` is false for 0/False
            double d = r - mu;
            double as = asinh(d / sig);
            j = exp(-0.5 * sq(goff + as)) / sqrt(d * d + sig * sig);
            break;
        }
    }
    if (has_g) j *= pow_small_int(g, specind);
    if (has_s) j *= pow_small_int(sinthb, 1.0 + specind);
    return j;
}

// ---- one image order at one pixel ------------------------------------------------------------------------------------
struct OrderOut {
    double I, Q, U, g, r, sinthb, Ir, Nmax, lp, Ie;
};

// Crossing bookkeeping of r_equatorial (kgeo/equatorial_lensing.py:158-180)
template <class W>
KG_HD void crossing_time(const W& w, const Observer& o, double beta, int mbar, double* Ir, double* Nmax) {
    double F0 = w(T_F0), K = w(T_KC), ww = w(T_W), tmax = w(S_TAUTOT);
    bool first = o.north ? (beta <= 0.0) : (beta >= 0.0);
    double heav = o.north ? ((beta > 0.0) ? 1.0 : 0.0) : ((beta < 0.0) ? 1.0 : 0.0);   // np.heaviside(+-beta, 0)
    double m = (double)mbar + heav;
    double sF = (o.north == (first ? 1 : 0)) ? -1.0 : 1.0;
    double v = floor((tmax * ww + sF * F0) / (2.0 * K));
    *Nmax = first ? v : v - 1.0;
    *Ir = (2.0 * m * K - sF * F0) / ww;
}

// Geometry of one order at one pixel: crossing Mino time, Nmax and crossing radius (r_equatorial, kgeo/equatorial_lensing.py:85-196).
// This is all of an order that needs the ray constants; the emission physics below needs only (r_s, Ir, Nmax) and five per-ray
// scalars, so the image kernel runs the geometry of ALL orders first and the ray constants are dead during the physics.
struct OrderGeom {
    double r, Ir, Nmax;
};
template <class W>
KG_HD OrderGeom image_geometry(const W& w, const Observer& o, int flags, double beta, int mbar) {
    OrderGeom gq{-1.0, -1.0, -2.0};
    if (!(w(S_ETA) > 0.0)) return gq;   // vortical / eta == 0: r_s = Ir = -1, Nmax = -2 (equatorial_lensing.py:119-123)
    double Ir, Nmax;
    crossing_time(w, o, beta, mbar, &Ir, &Nmax);
    double rs, d0, d1, d2;
    radial_eval<W, false>(w, o, flags, Ir, false, &rs, &d0, &d1, &d2);
    if (Nmax < (double)mbar) rs = 0.0;   // :188
    gq.r = rs;
    gq.Ir = Ir;
    gq.Nmax = Nmax;
    return gq;
}

// Emission of one order at one pixel from its crossing (equatorial_images.py:124-264 Iobs and helpers)
// POL = opt.polarization as a compile-time constant: the unpolarized kernel does not carry the tetrad / Penrose-Walker code's
// registers (two instantiations of the image kernel)
template <bool POL>
KG_HD OrderOut image_physics(const Observer& o, int flags, const kgeo_models& M, const kgeo_image_opts& opt, double alpha, double beta,
                             double lam, double eta, double tau_tot, int mbar, const OrderGeom& gq) {
    OrderOut out{0.0, 0.0, 0.0, 0.0, gq.r, 0.0, gq.Ir, gq.Nmax, 0.0, 0.0};
    if (!(eta > 0.0)) return out;
    const double rs = gq.r, Ir = gq.Ir, Nmax = gq.Nmax;
    bool dead = (Nmax < (double)mbar) || (Nmax == -1.0) || (rs <= o.rplus);   // equatorial_images.py:175
    if (M.emiscut_in > 0.0) dead = dead || (rs < M.emiscut_in);
    if (M.emiscut_out > 0.0) dead = dead || (rs > M.emiscut_out);
    if (dead) return out;   // a NaN radius is not masked (comparisons with NaN are False) and propagates, as in the reference
    const double a = o.a, a2 = o.a2, r = rs, r2 = rs * rs;
    double kr = ((!(flags & F_OUTSIDE)) || (Ir <= 0.5 * tau_tot)) ? 1.0 : -1.0;        // :266-288
    double par = (mbar & 1) ? -1.0 : 1.0;
    double kth = o.north ? -par : par;                                                      // :290-305
    // the models at the crossing radius and, for the thermal emissivity's field normalisation (:221-223), at its radius R_b
    // first: one rolled loop over one copy of the model code
    const bool thermal = (M.emissivity_type == KGEO_EM_THERMAL);
    const bool need_B = POL || thermal;
    FluidAt fl_at;
    double B0 = 1.0;
#pragma unroll 1
    for (int pass = thermal ? 1 : 0; pass >= 0; --pass) {
        const double rr = pass ? M.emissivity_p[0] : r;
        fl_at = fluid_at(M, a, rr, need_B);
        if (pass) B0 = sqrt(bsq_of(eq_metric(a, rr), fl_at));
    }
    const U4 u = fl_at.u;
    EqMetric gm = eq_metric(a, r);
    double Delta = gm.Delta;
    double Rpot = sq(r2 + a2 - a * lam) - Delta * (eta + sq(lam - a));
    double TH = eta;   // + a2 cos^2 - lam^2 cot^2 vanish in the plane
    double sqR = sqrt(Rpot), sqTH = sqrt(TH);
    double g = 1.0 / (u.u0 - lam * u.u3 - kth * u.u2 * sqTH - kr * u.u1 * sqR / Delta);   // :332
    out.g = g;
    double sinthb, bsq, c2chi = 0.0, s2chi = 0.0;
    if (POL) {   // calc_polquantities :336-487 + calc_evpa :524-544
        double l0 = gm.g00 * u.u0 + gm.g03 * u.u3, l1 = gm.g11 * u.u1, l2 = gm.g22 * u.u2, l3 = gm.g33 * u.u3 + gm.g03 * u.u0;
        double Nr = sqrt(-gm.g11 * (l0 * u.u0 + l3 * u.u3) * (1.0 + l2 * u.u2));
        double Nth = sqrt(gm.g22 * (1.0 + l2 * u.u2));
        double Nph = sqrt(-Delta * (l0 * u.u0 + l3 * u.u3));
        double ex[4] = {l1 * u.u0 / Nr, -(l0 * u.u0 + l3 * u.u3) / Nr, 0.0, l1 * u.u3 / Nr};
        double ey[4] = {l2 * u.u0 / Nth, l2 * u.u1 / Nth, (1.0 + l2 * u.u2) / Nth, l2 * u.u3 / Nth};
        double ez[4] = {l3 / Nph, 0.0, 0.0, -l0 / Nph};
        const double B1 = fl_at.B1, B2 = fl_at.B2, B3 = fl_at.B3;
        double b0 = B1 * l1 + B2 * l2 + B3 * l3;
        double b1 = (B1 + b0 * u.u1) / u.u0, b2 = (B2 + b0 * u.u2) / u.u0, b3 = (B3 + b0 * u.u3) / u.u0;
        double bl[4] = {gm.g00 * b0 + gm.g03 * b3, gm.g11 * b1, gm.g22 * b2, gm.g33 * b3 + gm.g03 * b0};
        double Bx = ex[0] * bl[0] + ex[1] * bl[1] + ex[2] * bl[2] + ex[3] * bl[3];
        double By = ey[0] * bl[0] + ey[1] * bl[1] + ey[2] * bl[2] + ey[3] * bl[3];
        double Bz = ez[0] * bl[0] + ez[1] * bl[1] + ez[2] * bl[2] + ez[3] * bl[3];
        double Bmag = sqrt(Bx * Bx + By * By + Bz * Bz);
        bsq = Bmag * Bmag;
        double g00u = -(r2 + a2 + 2.0 * r * a2 / gm.Sigma) / Delta;
        double g11u = Delta / gm.Sigma, g22u = 1.0 / gm.Sigma;
        double g33u = (Delta - a2) / (gm.Sigma * Delta);
        double g03u = -(2.0 * a * r) / (gm.Sigma * Delta);
        double kl[4] = {-1.0, kr * sqR / Delta, kth * sqTH, lam};
        double k0 = g00u * kl[0] + g03u * kl[3], k1 = g11u * kl[1], k2 = g22u * kl[2], k3 = g33u * kl[3] + g03u * kl[0];
        double kx = ex[0] * kl[0] + ex[1] * kl[1] + ex[2] * kl[2] + ex[3] * kl[3];
        double ky = ey[0] * kl[0] + ey[1] * kl[1] + ey[2] * kl[2] + ey[3] * kl[3];
        double kz = ez[0] * kl[0] + ez[1] * kl[1] + ez[2] * kl[2] + ez[3] * kl[3];
        double kmag = sqrt(kx * kx + ky * ky + kz * kz);
        double fx = (ky * Bz - kz * By) / kmag, fy = (kz * Bx - kx * Bz) / kmag, fz = (kx * By - ky * Bx) / kmag;
        sinthb = sqrt(fx * fx + fy * fy + fz * fz) / Bmag;
        double f0 = ex[0] * fx + ey[0] * fy + ez[0] * fz;
        double f1 = ex[1] * fx + ey[1] * fy + ez[1] * fz;
        double f2 = ex[2] * fx + ey[2] * fy + ez[2] * fz;
        double f3 = ex[3] * fx + ey[3] * fy + ez[3] * fz;
        double A = k0 * f1 - k1 * f0 + a * (k1 * f3 - k3 * f1);
        double B = ((r2 + a2) * (k3 * f2 - k2 * f3) - a * (k0 * f2 - k2 * f0));
        // kappa = (r - i a cos th)(A - iB) with cos th = 6e-17 kept as in the reference
        double ka1 = r * A - (a * kCosHalfPi) * B;
        double ka2 = -(r * B) - (a * kCosHalfPi) * A;
        double mu = -(alpha + a * o.sth);
        double den = (beta * beta + mu * mu) * (ka1 * ka1 + ka2 * ka2);
        double p1 = beta * ka1 + mu * ka2, p2 = mu * ka1 - beta * ka2;
        c2chi = (p1 * p1 - p2 * p2) / den;
        s2chi = (2.0 * p1 * p2) / den;
    } else {
        sinthb = kHalfPi * 0.5;   // SINTHB = pi/4, equatorial_images.py:20,206
        bsq = thermal ? bsq_of(gm, fl_at) : 0.0;   // only the thermal emissivity reads the field strength
    }
    out.sinthb = sinthb;
    double Bmag = sqrt(bsq);
    if (thermal) Bmag = (Bmag / B0) * M.emissivity_p[3];   // :221-223
    double je = emissivity_jrest(M, a, r, true, g, true, sinthb, Bmag, opt.nu_obs, opt.specind);
    double I;
    if (opt.pathlength) {   // calc_pathlength_equatorial :489-522
        double kyhat = -(kth * sqTH) / r;
        double llp = (opt.diskangle * r) * ((1.0 / g) / kyhat);
        je *= fabs(llp);
        I = (g * g * g) * je;
        out.Ie = je;
        out.lp = llp;
    } else {
        I = (g * g) * je;
    }
    out.I = I;
    if (POL) {
        out.Q = c2chi * I;
        out.U = s2chi * I;
    }
    return out;
}

// one order, geometry and emission together (host-side checks; the image kernel calls the two halves separately)
template <class W>
KG_HD OrderOut image_order(const W& w, const Observer& o, int flags, const kgeo_models& M, const kgeo_image_opts& opt,
                           double alpha, double beta, int mbar) {
    const OrderGeom gq = image_geometry(w, o, flags, beta, mbar);
    return opt.polarization ? image_physics<true>(o, flags, M, opt, alpha, beta, w(S_LAM), w(S_ETA), w(S_TAUTOT), mbar, gq)
                            : image_physics<false>(o, flags, M, opt, alpha, beta, w(S_LAM), w(S_ETA), w(S_TAUTOT), mbar, gq);
}

}  // namespace kg
