// DEVELOPER-ONLY harness: host compile of kg_offeq.cuh (off-equatorial jet physics) for CPU checks against the oracle.
#include "../../kgeo_b200/csrc/kg_offeq.cuh"
using namespace kg;
extern "C" {
void hs_offeq(int field, double C, double pval, int usemono, const double* fr_x, const double* fr_f, long long fr_n, double a,
              double th_o, int n, const double* alpha, const double* beta, const double* r, const double* th, const double* kr,
              const double* kth, double nu, const double* density, double specind, double* I, double* Q, double* U, double* sinthb,
              double* g, double* B, double* u) {
    JetSpec J{field, C, pval, usemono, fr_x, fr_f, fr_n};
    for (int i = 0; i < n; ++i) {
        StokesOut s = iobs_off_point(J, a, th_o, alpha[i], beta[i], r[i], th[i], kr[i], kth[i], nu, density[i], specind);
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
    for (int i = 0; i < n; ++i) rho[i] = jet_density(J, a, r[i], th[i], nu, sigma, r0, norm, psitarget);
}
double hs_power_source_integral(double r, double r0, double a, double sigma, double psi, double pval) {
    return power_source_integral(r, r0, a, sigma, psi, pval);
}
}
