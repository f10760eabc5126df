// Field-line crossings of a geodesic (kgeo/solver.py): the stream functions psi(r, theta) of the BZ jet models, the Mino
// time up to which crossings are searched, and the crossing-count bookkeeping of `getneq`.  Per-sample / per-ray device
// functions; the kernels and the secant driver are in kgeo_b200.cu.
#pragma once
#include "kg_geodesic.cuh"

namespace kg {

enum PsiModel { PSI_PARA = 0, PSI_MONO = 1, PSI_POWER = 2, PSI_NONE = 3 };

// psifunc (kgeo/solver.py:13-37)
KG_HD double psi_value(int model, double pval, double a, double r, double th) {
    const double c = fabs(cos(th));
    if (model == PSI_PARA) {
        const double rp = 1.0 + sqrt(1.0 - a * a);
        return r * (1.0 - c) + rp * (1.0 + c) * (1.0 - log(1.0 + c)) - 2.0 * rp * (1.0 - 0.6931471805599453);
    }
    if (model == PSI_MONO) return 1.0 - c;
    if (model == PSI_POWER) return pow(r, pval) * (1.0 - c);
    return 0.0;
}

KG_HD double np_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : ((x == 0.0) ? 0.0 : x)); }   // NaN stays NaN
KG_HD double np_heaviside0(double x) { return (x > 0.0) ? 1.0 : ((x <= 0.0) ? 0.0 : x); }                    // heaviside(x, 0)

// get_maxtau_forwardjet (kgeo/solver.py:57-88): the Mino time of the neqmax-th equatorial crossing (Eq. 81 of Gralla &
// Lupsasca's lensing paper), capped by MAXTAUFRAC * tau_tot; neqmax < 0 = no limit
template <class W>
KG_HD double maxtau_forwardjet(const W& w, int flags, double beta, int neqmax, double maxtaufrac) {
    const double tt2 = w(S_TAUTOT) * maxtaufrac;
    if (neqmax < 0) return tt2;
    if (!(w(S_ETA) > 0.0)) return tt2;                 // vortical: taumax[eta <= 0] = tau_tot2
    const double mmax = (double)(neqmax - 1) + np_heaviside0(beta);
    const double taumax = (2.0 * mmax * w(T_KC) - np_sign(beta) * w(T_F0)) / w(T_W);
    if (taumax != taumax || tt2 != tt2) return kg_nan();   // np.min propagates NaN
    return (taumax < tt2) ? taumax : tt2;
}

// getneq (kgeo/solver.py:40-54): ceil(nan_to_num((sqrt(-u_- a^2) tau + sign(beta) F0) / 2K - heaviside(beta, 0)))
template <class W>
KG_HD double neq_of_tau(const W& w, int flags, double beta, double tau) {
    if (!(flags & F_TH_ORD)) return 0.0;               // u_- > 0: sqrt(-u_- a^2) is NaN -> nan_to_num -> 0
    double mbar = (w(T_W) * tau + np_sign(beta) * w(T_F0)) / (2.0 * w(T_KC)) - np_heaviside0(beta);
    if (mbar != mbar) mbar = 0.0;
    else if (mbar > 1.7976931348623157e308) mbar = 1.7976931348623157e308;
    else if (mbar < -1.7976931348623157e308) mbar = -1.7976931348623157e308;
    return ceil(mbar);
}

// psi(r(tau), theta(tau)) - psitarget of one ray (the objective of findroot's Newton solve, kgeo/solver.py:163-175)
template <class W>
KG_HD double psi_residual(const W& w, const Observer& o, int flags, int model, double pval, double psitarget, double tau,
                          double* r_out, double* th_out) {
    const Point p = point_eval(w, o, flags, tau, false);
    if (r_out) *r_out = p.r;
    if (th_out) *th_out = p.th;
    return psi_value(model, pval, o.a, p.r, p.th) - psitarget;
}

}  // namespace kg
