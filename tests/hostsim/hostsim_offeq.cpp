// DEVELOPER-ONLY harness: host compile of kg_offeq.cuh (off-equatorial jet physics) for CPU checks against the oracle.
#include <cmath>
#include "../../kgeo_b200/csrc/kg_offeq.cuh"
using namespace kg;
extern "C" {
void hs_offeq(int field, double C, double pval, int usemono, const double* fr_x, const double* fr_f, long long fr_n, double a,
              double th_o, int n, const double* alpha, const double* beta, const double* r, const double* th, const double* kr,
              const double* kth, double nu, const double* density, double specind, double* I, double* Q, double* U, double* sinthb,
              double* g, double* B, double* u) {
    JetSpec J{field, C, pval, usemono, fr_x, fr_f, fr_n};
    for (int i = 0; i < n; ++i) {
        StokesOut s = iobs_off_point(J, a, th_o, alpha[i], beta[i], r[i], th[i], kr[i], kth[i], jet_boost_numeric(nu), density[i], specind);
        I[i] = s.I; Q[i] = s.Q; U[i] = s.U; sinthb[i] = s.sinthb; g[i] = s.g;
        double om;
        jet_field(J, a, r[i], th[i], &B[4 * i], &B[4 * i + 1], &B[4 * i + 2], &om);
        B[4 * i + 3] = om;
        U4 v = jet_driftframe(J, a, r[i], th[i], nu);
        u[4 * i] = v.u0; u[4 * i + 1] = v.u1; u[4 * i + 2] = v.u2; u[4 * i + 3] = v.u3;
    }
}
void hs_jet_density(int field, double C, double pval, int usemono, const double* fr_x, const double* fr_f, long long fr_n, double a,
                    int n, const double* r, const double* th, double nu, double sigma, double r0, double norm, double psitarget,
                    double* rho) {
    JetSpec J{field, C, pval, usemono, fr_x, fr_f, fr_n};
    for (int i = 0; i < n; ++i) rho[i] = jet_density(J, a, r[i], th[i], jet_boost_numeric(nu), sigma, r0, norm, psitarget);
}
// the whole per-crossing pipeline of kgeo_jet_stokes on the host, in the kernels' order: gammamax pre-pass (gamma, gamma_eff, first
// argmin of |nan_to_num(gamma_eff)|), then density + Stokes parameters with the FF boost constants (global / per crossing order)
int hs_jet_stokes(int field, double C, double pval, int usemono, const double* fr_x, const double* fr_f, long long fr_n, double a,
                  double th_o, int npix, int nguess, const double* alpha, const double* beta, const double* r, const double* th,
                  const double* kr, const double* kth, double nu, int nu_ff, const double* ff_global, const double* ff_rows,
                  double gammamax, double sigma, double r0, double norm, double psitarget, double specind, double* I, double* Q,
                  double* U, double* sinthb, double* dens) {
    JetSpec J{field, C, pval, usemono, fr_x, fr_f, fr_n};
    JetBoost boost = jet_boost_numeric(nu);
    if (nu_ff) { boost.ff = 1; boost.r0 = ff_global[0]; boost.omega = ff_global[1]; boost.Aconst = ff_global[2]; }
    boost.gammamax = gammamax > 0.0 ? gammamax : 0.0;
    const int n = npix * nguess;
    if (boost.gammamax > 0.0) {
        double best = KG_INF;
        int arg = 0;
        double gbest = 1.0, ebest = 1.0;
        for (int e = 0; e < n; ++e) {
            JetBoost b = boost; b.gscale = 1.0;
            DriftOut d = jet_driftframe_boost(J, a, r[e], th[e], b);
            double x = d.gammaeff, key = (x != x) ? KG_INF : (std::isinf(x) ? 1.7976931348623157e308 : fabs(x));
            if (key < best || e == 0) { if (key < best || e == 0) { best = key; arg = e; gbest = d.gamma; ebest = d.gammaeff; } }
        }
        (void)arg;
        boost.gscale = gbest / ebest;
    }
    int bad = 0;
    for (int e = 0; e < n; ++e) {
        JetBoost bd = boost; bd.gammamax = 0.0; bd.gscale = 1.0;
        if (nu_ff) { const int m = e / npix; bd.r0 = ff_rows[3 * m]; bd.omega = ff_rows[3 * m + 1]; bd.Aconst = ff_rows[3 * m + 2]; }
        double nu_d = 0.0;
        const double rho = jet_density(J, a, r[e], th[e], bd, sigma, r0, norm, psitarget, &nu_d);
        dens[e] = rho;
        StokesOut s = iobs_off_point(J, a, th_o, alpha[e % npix], beta[e % npix], r[e], th[e], kr[e], kth[e], boost, rho, specind);
        if (s.nu > 1.0 || s.nu < -1.0 || nu_d > 1.0 || nu_d < -1.0) bad = 1;
        const bool none = (r[e] == 0.0);
        auto n2n = [](double x) { return (x != x) ? 0.0 : (x > 1.7976931348623157e308 ? 1.7976931348623157e308 : (x < -1.7976931348623157e308 ? -1.7976931348623157e308 : x)); };
        I[e] = none ? 0.0 : n2n(s.I); Q[e] = none ? 0.0 : n2n(s.Q); U[e] = none ? 0.0 : n2n(s.U); sinthb[e] = n2n(s.sinthb);
    }
    return bad;
}
double hs_power_source_integral(double r, double r0, double a, double sigma, double psi, double pval) {
    return power_source_integral(r, r0, a, sigma, psi, pval);
}
}
