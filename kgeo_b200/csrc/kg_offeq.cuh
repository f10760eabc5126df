// Off-equatorial emission physics of the jet models (kgeo/off_eq.py Iobs_off and what it calls at theta != pi/2): the Kerr
// metric, the BZ monopole / power-law fields with their field-line speed, the E x B drift frame, redshift, fluid-frame
// tetrad, Penrose-Walker transport -- per-crossing device functions for the rest of SURVEY 8 f4.
//
// Used by jet_stokes_kernel (kgeo_b200.cu).  tests/hostsim/check_offeq.py also compiles this header for the host and compares every
// output with the reference-pinned CPU checker on the crossings of tests/golden/jet_pipeline.npz.  The equatorial image kernel
// keeps its own plane-specialised versions (kg_image.cuh).
#pragma once
#include "kg_image.cuh"

namespace kg {

struct KerrMetric {
    double Delta, Sigma, g00, g11, g22, g33, g03, sth, cth, sth2, cth2;
};
KG_HD KerrMetric kerr_metric(double a, double r, double th) {   // kgeo/velocities.py:102-114
    KerrMetric g;
    const double a2 = a * a, r2 = r * r;
    g.sth = sin(th);
    g.cth = cos(th);
    g.sth2 = g.sth * g.sth;
    g.cth2 = g.cth * g.cth;
    g.Delta = r2 - 2.0 * r + a2;
    g.Sigma = r2 + a2 * g.cth2;
    g.g00 = -(1.0 - 2.0 * r / g.Sigma);
    g.g11 = g.Sigma / g.Delta;
    g.g22 = g.Sigma;
    g.g33 = (r2 + a2 + 2.0 * r * a2 * g.sth2 / g.Sigma) * g.sth2;
    g.g03 = -2.0 * r * a * g.sth2 / g.Sigma;
    return g;
}

enum JetField { JET_MONO = 0, JET_POWER = 1 };
struct JetSpec {
    int field;            // JET_MONO: Bfield('bz_monopole', C) ; JET_POWER: Bfield('power', p, C, usemono)
    double C, pval;
    int usemono;
    const double* fr_x;   // f(r) table of the monopole (see kgeo_b200/bfields.py::bz_fr_table)
    const double* fr_f;
    long long fr_n;
};

// lab-frame B^i and the field-line angular speed at (r, theta): Bfield_BZmonopole (kgeo/bfields.py:545-592, first-order
// branch) and Bfield_power (:676-709) with omega_BZpower (:618-625)
KG_HD void jet_field(const JetSpec& J, double a, double r, double th, double* B1, double* B2, double* B3, double* omega) {
    const double a2 = a * a, r2 = r * r;
    const double sth = sin(th), cth = cos(th), sth2 = sth * sth, cth2 = cth * cth;
    const double Delta = r2 - 2.0 * r + a2, Sigma = r2 + a2 * cth2, gdet = sth * Sigma;
    if (J.field == JET_MONO) {
        double fr, frp;
        linear_spline(J.fr_x, J.fr_f, J.fr_n, r, &fr, &frp);
        if (r > J.fr_x[J.fr_n - 1]) {
            fr = 1.0 / (4.0 * r);
            frp = -1.0 / (4.0 * r2);
        }
        const double dphidtheta = J.C * sth * (1.0 + 0.5 * a2 * fr * (1.0 + 3.0 * cos(2.0 * th)));
        const double dphidr = J.C * a2 * sth2 * cth * frp;
        const double f2 = (6.0 * kPi * kPi - 49.0) / 72.0;
        const double omega2 = 1.0 / 32.0 - sth2 * (4.0 * f2 - 1.0) / 64.0;
        const double I = -1.0 * J.C * 2.0 * kPi * sth2 * (a / 8.0 + (a * a * a) * (omega2 + 0.25 * fr * cth2));
        *B1 = dphidtheta / gdet;
        *B2 = -dphidr / gdet;
        *B3 = I / (2.0 * kPi * Delta * sth2);
        const double spin = fabs(a);
        const double Om = spin / 8.0 + (spin * spin * spin) * omega2;
        *omega = (a < 0.0) ? -Om : Om;
        return;
    }
    const double abscth = fabs(cth), sg = sgn0(cth);
    const double rp = 1.0 + sqrt(1.0 - a2);
    const double rpw = pow(r, J.pval);
    const double psi = rpw * (1.0 - abscth);
    const double dpsidtheta = sg * sth * rpw;
    const double dpsidr = J.pval * psi / r;
    double Om;
    if (J.usemono || J.pval == 0.0) {
        Om = a / 8.0;
    } else {
        const double cthh = 1.0 - psi * pow(rp, -J.pval);
        Om = a / (4.0 + 8.0 / (1.0 + cthh));
    }
    const double I = (J.pval > 0.0) ? (-4.0 * kPi * psi * Om * sg) : (-2.0 * kPi * psi * (2.0 - psi) * Om * sg);
    *B1 = J.C * dpsidtheta / gdet;
    *B2 = -J.C * dpsidr / gdet;
    *B3 = I / (2.0 * kPi * Delta * sth2);
    *omega = Om;
}

// How u_driftframe picks its parallel boost and caps the Lorentz factor (kgeo/velocities.py:491-516,593-611)
struct JetBoost {
    int ff;                     // 1: nu_parallel = 'FF', the energy-conserving boost of getnu_cons (:474-486)
    double nu;                  // numeric nu_parallel (ff == 0)
    double r0, omega, Aconst;   // FF: stagnation radius, angular speed and E - Omega_F L (ff_boost.getEco) of the field line
    double gammamax, gscale;    // gammamax > 0: gamma_eff = (gammamax^-2 + gamma^-2)^(-1/2) * gscale (gscale = gamma / gamma_eff at
                                // the element of smallest gamma_eff: "ensure gamma > 1 always")
};
KG_HD JetBoost jet_boost_numeric(double nu) { return JetBoost{0, nu, 0.0, 0.0, 0.0, 0.0, 1.0}; }
struct DriftOut {
    U4 u;
    double nu, gamma, gammaeff;   // the boost used, the force-free Lorentz factor and (gammamax > 0) the capped one before scaling
};

// drift-frame four-velocity at (r, theta) (kgeo/velocities.py:487-623)
KG_HD DriftOut jet_driftframe_boost(const JetSpec& J, double a, double r, double th, const JetBoost& b) {
    const double a2 = a * a, r2 = r * r;
    const double sth = sin(th), cth = cos(th), sth2 = sth * sth, cth2 = cth * cth;
    const double Delta = r2 - 2.0 * r + a2, Sigma = r2 + a2 * cth2;
    const double Xi = sq(r2 + a2) - Delta * a2 * sth2;
    const double omegaz = 2.0 * a * r / Xi;
    const double gdet = Sigma * sth;
    const double g11 = Sigma / Delta, g22 = Sigma, g33 = Xi * sth2 / Sigma;
    const double alpha = sqrt(Delta * Sigma / Xi);
    const double eta3 = 2.0 * a * r / sqrt(Delta * Sigma * Xi);
    double B1, B2, B3, omegaf;
    jet_field(J, a, r, th, &B1, &B2, &B3, &omegaf);
    const double E1 = (omegaf - omegaz) * Xi * sth * B2 / Sigma;
    const double E2 = -(omegaf - omegaz) * Xi * sth * B1 / (Sigma * Delta);
    const double Bsq = g11 * B1 * B1 + g22 * B2 * B2 + g33 * B3 * B3;
    const double Esq = g11 * E1 * E1 + g22 * E2 * E2;
    const double B1c = g11 * B1, B2c = g22 * B2, B3c = g33 * B3, E1c = g11 * E1, E2c = g22 * E2;
    const double pre = alpha / (Bsq * gdet);
    const double vperp1 = pre * (E2c * B3c);
    const double vperp2 = pre * (-B3c * E1c);
    const double vperp3 = pre * (E1c * B2c - B1c * E2c);
    const double vpar_max = sqrt(1.0 - Esq / Bsq), sB = sqrt(Bsq);
    DriftOut out;
    double nu = b.nu;
    if (b.ff) {   // getnu_cons: the boost that conserves E - Omega_F L along the field line, from the nu = 0 frame's quantities
        const KerrMetric gm = kerr_metric(a, r, th);
        const double gfac = gm.g03 + gm.g33 * b.omega;
        const double gammap = 1.0 / vpar_max, Bhatphi = B3 / sB;
        const double ffunc = gammap * (alpha - gfac * vperp3);
        const double bred = Bhatphi * gfac;
        const double A = b.Aconst;
        const double nunum = ffunc * bred + sgn0(cos(th)) * sgn0(r - b.r0) * A * sqrt(A * A - ffunc * ffunc + bred * bred);
        nu = nunum / (A * A + bred * bred);
    }
    out.nu = nu;
    double v1 = vperp1 + nu * vpar_max * B1 / sB;
    double v2 = vperp2 + nu * vpar_max * B2 / sB;
    double v3 = vperp3 + nu * vpar_max * B3 / sB;
    const double vsq = g11 * v1 * v1 + g22 * v2 * v2 + g33 * v3 * v3;
    double gamma = 1.0 / sqrt(1.0 - vsq);
    out.gamma = gamma;
    out.gammaeff = gamma;
    if (b.gammamax > 0.0) {   // approximate MHD gamma: gamma_FF and gamma_max summed "in series" (:593-611, eps < 0 branch)
        const double gammaeff = pow(1.0 / (b.gammamax * b.gammamax) + 1.0 / (gamma * gamma), -0.5);
        out.gammaeff = gammaeff;
        const double gammaeff0 = gammaeff * b.gscale;
        const double vsqeff = 1.0 - 1.0 / (gammaeff0 * gammaeff0);
        const double sc = sqrt(vsqeff / vsq);
        v1 *= sc; v2 *= sc; v3 *= sc;
        gamma = gammaeff0;
    }
    out.u = U4{gamma / alpha, gamma * v1, gamma * v2, gamma * (v3 + eta3)};
    return out;
}
KG_HD U4 jet_driftframe(const JetSpec& J, double a, double r, double th, double nu_parallel) {
    return jet_driftframe_boost(J, a, r, th, jet_boost_numeric(nu_parallel)).u;
}

struct StokesOut {
    double I, Q, U, sinthb, g, nu;
};

// One crossing of Iobs_off as getstokes calls it (kgeo/off_eq.py:15-121: Emissivity('constant'), polarization=True):
// redshift (equatorial_images.py:307-334), tetrad + fluid-frame field + path length + Penrose-Walker kappa (:336-487), EVPA
// (:524-544), I = g^(3+specind) l_p density b^(1+specind) sin^(1+specind) theta_b, zero inside the horizon
KG_HD StokesOut iobs_off_point(const JetSpec& J, double a, double th_o, double alpha, double beta, double r, double th,
                               double kr_sign, double kth_sign, const JetBoost& boost, double density, double specind) {
    StokesOut out{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double a2 = a * a, r2 = r * r;
    const double sth_o = sin(th_o), cth_o = cos(th_o);
    const double lam = -alpha * sth_o;
    const double eta = (alpha * alpha - a2) * cth_o * cth_o + beta * beta;
    const KerrMetric gm = kerr_metric(a, r, th);
    const double Delta = gm.Delta, Sigma = gm.Sigma, sth2 = gm.sth2, cth2 = gm.cth2;
    const DriftOut dr = jet_driftframe_boost(J, a, r, th, boost);
    const U4 u = dr.u;
    out.nu = dr.nu;
    const double Rpot = sq(r2 + a2 - a * lam) - Delta * (eta + sq(lam - a));
    const double TH = eta + a2 * cth2 - lam * lam * cth2 / sth2;
    const double sqR = sqrt(Rpot), sqTH = sqrt(TH);
    const double g = 1.0 / (u.u0 - lam * u.u3 - sgn0(kth_sign) * u.u2 * sqTH - sgn0(kr_sign) * u.u1 * sqR / Delta);
    out.g = g;
    // covariant velocity and tetrad (kgeo/velocities.py:98-176)
    const double l0 = gm.g00 * u.u0 + gm.g03 * u.u3, l1 = gm.g11 * u.u1, l2 = gm.g22 * u.u2, l3 = gm.g33 * u.u3 + gm.g03 * u.u0;
    const double Nr = sqrt(-gm.g11 * (l0 * u.u0 + l3 * u.u3) * (1.0 + l2 * u.u2));
    const double Nth = sqrt(gm.g22 * (1.0 + l2 * u.u2));
    const double Nph = sqrt(-Delta * sth2 * (l0 * u.u0 + l3 * u.u3));
    const double ex[4] = {l1 * u.u0 / Nr, -(l0 * u.u0 + l3 * u.u3) / Nr, 0.0, l1 * u.u3 / Nr};
    const double ey[4] = {l2 * u.u0 / Nth, l2 * u.u1 / Nth, (1.0 + l2 * u.u2) / Nth, l2 * u.u3 / Nth};
    const double ez[4] = {l3 / Nph, 0.0, 0.0, -l0 / Nph};
    double B1, B2, B3, om;
    jet_field(J, a, r, th, &B1, &B2, &B3, &om);
    const double gdet = Sigma * sqrt(sth2);
    const double nl[4] = {0.0, -B2 * gdet, B1 * gdet, 0.0};
    const double b0 = B1 * l1 + B2 * l2 + B3 * l3;
    const double b1 = (B1 + b0 * u.u1) / u.u0, b2 = (B2 + b0 * u.u2) / u.u0, b3 = (B3 + b0 * u.u3) / u.u0;
    const double bl[4] = {gm.g00 * b0 + gm.g03 * b3, gm.g11 * b1, gm.g22 * b2, gm.g33 * b3 + gm.g03 * b0};
    auto dot = [](const double* e, const double* v) { return e[0] * v[0] + e[1] * v[1] + e[2] * v[2] + e[3] * v[3]; };
    const double Bx = dot(ex, bl), By = dot(ey, bl), Bz = dot(ez, bl);
    const double nx = dot(ex, nl), ny = dot(ey, nl), nz = dot(ez, nl);
    const double Bmag = sqrt(Bx * Bx + By * By + Bz * Bz), bsq = Bmag * Bmag;
    const double nmag = sqrt(nx * nx + ny * ny + nz * nz);
    const double g00u = -(r2 + a2 + 2.0 * r * a2 * sth2 / Sigma) / Delta, g11u = Delta / Sigma, g22u = 1.0 / Sigma;
    const double g33u = (Delta - a2 * sth2) / (Sigma * Delta * sth2), g03u = -(2.0 * a * r) / (Sigma * Delta);
    const double kl[4] = {-1.0, kr_sign * sqR / Delta, kth_sign * sqTH, lam};
    const double k0 = g00u * kl[0] + g03u * kl[3], k1 = g11u * kl[1], k2 = g22u * kl[2], k3 = g33u * kl[3] + g03u * kl[0];
    const double kx = dot(ex, kl), ky = dot(ey, kl), kz = dot(ez, kl);
    const double kmag = sqrt(kx * kx + ky * ky + kz * kz);
    const double kdotn = kx * nx / nmag + ky * ny / nmag + kz * nz / nmag;
    const double pathlength = fabs(kmag / kdotn);
    const double fx = (ky * Bz - kz * By) / kmag, fy = (kz * Bx - kx * Bz) / kmag, fz = (kx * By - ky * Bx) / kmag;
    const double sinthb = sqrt(fx * fx + fy * fy + fz * fz) / Bmag;
    out.sinthb = sinthb;
    const double f0 = ex[0] * fx + ey[0] * fy + ez[0] * fz, f1 = ex[1] * fx + ey[1] * fy + ez[1] * fz;
    const double f2 = ex[2] * fx + ey[2] * fy + ez[2] * fz, f3 = ex[3] * fx + ey[3] * fy + ez[3] * fz;
    const double A = k0 * f1 - k1 * f0 + a * sth2 * (k1 * f3 - k3 * f1);
    const double B = ((r2 + a2) * (k3 * f2 - k2 * f3) - a * (k0 * f2 - k2 * f0)) * gm.sth;
    const double ka1 = r * A - (a * gm.cth) * B;          // kappa = (r - i a cos th)(A - i B)
    const double ka2 = -(r * B) - (a * gm.cth) * A;
    const double mu = -(alpha + a * sth_o);
    const double den = (beta * beta + mu * mu) * (ka1 * ka1 + ka2 * ka2);
    const double p1 = beta * ka1 + mu * ka2, p2 = mu * ka1 - beta * ka2;
    const double c2chi = (p1 * p1 - p2 * p2) / den, s2chi = (2.0 * p1 * p2) / den;
    const double Iemis = density * pow(bsq, 0.5 * (1.0 + specind));
    double I = (g * g * g) * pow(g, specind) * pathlength * Iemis * pow(sinthb, 1.0 + specind);
    const double rh = 1.0 + sqrt(1.0 - a2);
    if (r <= rh) {
        out.I = 0.0; out.Q = 0.0; out.U = 0.0;
        return out;
    }
    out.I = I;
    out.Q = c2chi * I;
    out.U = s2chi * I;
    return out;
}

}  // namespace kg

// ---- mass loading and density of the jet models (kgeo/densityfuncs.py) ----------------------------------------------------
namespace kg {

// eta of the monopole jet: Gaussian source integrated in closed form (kgeo/densityfuncs.py:31-41)
KG_HD double eta_mono(double r, double th, double r0, double a, double sigma) {
    if (sigma == 0.0) return sgn0(r - r0);
    const double cth = cos(th), s2pi = 2.5066282746310002;   // sqrt(2 pi)
    const double d = r - r0;
    const double num1 = 2.0 * sigma * (2.0 * r0 - (r + r0) * exp(-d * d / (2.0 * sigma * sigma)));
    const double num2 = s2pi * erf(d / (1.4142135623730951 * sigma)) * (r0 * r0 + sigma * sigma + a * a * cth * cth);
    return (num1 + num2) / (4.0 * r0 * sigma + s2pi * sigma * sigma + s2pi * (r0 * r0 + a * a * cth * cth));
}

// integral of the power-law jet's source exp(-(x - r0)^2 / 2 sigma^2) (x^2 + a^2 (1 - psi / x^p)^2) from r0 to r
// (kgeo/densityfuncs.py:24-28; the reference uses QUADPACK at 1e-13): 16-point Gauss-Legendre on panels one sigma wide; the
// Gaussian makes everything beyond 40 sigma vanish in double precision
KG_HD double power_source_integral(double r, double r0, double a, double sigma, double psi, double pval) {
    const double xg[8] = {0.0950125098376374402, 0.2816035507792589132, 0.4580167776572273863, 0.6178762444026437484,
                          0.7554044083550030339, 0.8656312023878317439, 0.9445750230732325761, 0.9894009349916499326};
    const double wg[8] = {0.1894506104550684963, 0.1826034150449235888, 0.1691565193950025382, 0.1495959888165767321,
                          0.1246289712555338720, 0.0951585116824927848, 0.0622535239386478929, 0.0271524594117540949};
    const double dir = (r >= r0) ? 1.0 : -1.0;
    double span = fabs(r - r0);
    if (span > 40.0 * sigma) span = 40.0 * sigma;
    const int npanel = (int)ceil(span / sigma);
    double total = 0.0;
    for (int k = 0; k < npanel; ++k) {
        const double lo = k * sigma, hi = ((k + 1) * sigma < span) ? (k + 1) * sigma : span;
        const double c = 0.5 * (lo + hi), h = 0.5 * (hi - lo);
        double acc = 0.0;
        for (int i = 0; i < 8; ++i) {
            for (int s = -1; s <= 1; s += 2) {
                const double t = c + s * h * xg[i];           // distance from r0
                const double x = r0 + dir * t;
                const double q = 1.0 - psi / pow(x, pval);
                acc += wg[i] * exp(-t * t / (2.0 * sigma * sigma)) * (x * x + a * a * q * q);
            }
        }
        total += acc * h;
    }
    return dir * total;
}

// rho = eta B^r / u^r at a crossing reflected into the northern hemisphere (density_mono_all / density_power_all,
// kgeo/densityfuncs.py:44-79,186-219); r0 = stagnation radius, norm = source integral from r0 to infinity (power law)
KG_HD double jet_density(const JetSpec& J, double a, double r, double th, const JetBoost& boost, double sigma, double r0, double norm,
                         double psitarget, double* nu_used = nullptr) {
    const double thn = acos(fabs(cos(th)));
    double B1, B2, B3, om;
    jet_field(J, a, r, thn, &B1, &B2, &B3, &om);
    const DriftOut dr = jet_driftframe_boost(J, a, r, thn, boost);   // the density's velocity never caps gamma (gammamax = None there)
    const U4 u = dr.u;
    if (nu_used) *nu_used = dr.nu;
    double eta;
    if (J.field == JET_MONO) eta = eta_mono(r, thn, r0, a, sigma);
    else if (sigma == 0.0) eta = sgn0(r - r0);
    else eta = (r == 0.0) ? 0.0 : power_source_integral(r, r0, a, sigma, psitarget, J.pval) / norm;
    const double rho = eta * B1 / u.u1;
    return (rho == rho) ? rho : 0.0;                          // nan_to_num(nan = 0)
}

}  // namespace kg
