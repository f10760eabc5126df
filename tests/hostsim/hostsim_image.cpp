// DEVELOPER-ONLY harness (see hostsim_special.cpp): host compile of the device image/lensing math for algorithm checks.
#include "../../kgeo_b200/csrc/kg_image.cuh"
#include <vector>
using namespace kg;
extern "C" {
void hs_image(double a, double r_o, double th_o, const kgeo_models* M, const kgeo_image_opts* opt, int n, const double* alpha,
              const double* beta, double* I, double* Q, double* U, double* r, double* Ir, double* g, double* sinthb, double* nmax,
              double* npol, double* imax, double* lp, double* Ie) {
    Observer o = make_observer(a, r_o, th_o);
    std::vector<double> w(NW);
    for (int i = 0; i < n; ++i) {
        int fl;
        ray_setup<false, false>(o, alpha[i], beta[i], M_ANY_ETA_POS, w.data(), &fl);
        ArrayW W{w.data()};
        npol[i] = n_poloidal_at(W, w[S_TAUTOT]);
        imax[i] = (w[S_ETA] > 0.0) ? w[S_TAUTOT] : -1.0;
        double lastN = -2.0;
        for (int mbar = opt->mbar_min; mbar <= opt->mbar_max; ++mbar) {
            OrderOut v = image_order(W, o, fl, *M, *opt, alpha[i], beta[i], mbar);
            size_t q = (size_t)(mbar - opt->mbar_min) * n + i;
            I[q] = v.I; Q[q] = v.Q; U[q] = v.U; r[q] = v.r; Ir[q] = v.Ir; g[q] = v.g; sinthb[q] = v.sinthb; lp[q] = v.lp; Ie[q] = v.Ie;
            lastN = v.Nmax;
        }
        nmax[i] = lastN;
    }
}
void hs_requatorial(double a, double r_o, double th_o, int n, const double* alpha, const double* beta, int mlo, int mhi,
                    double* rs_o, double* Ir_o, double* Nmax_o, double* Imax_o, double* npol_o) {
    Observer o = make_observer(a, r_o, th_o);
    std::vector<double> w(NW);
    for (int i = 0; i < n; ++i) {
        int fl;
        ray_setup<false, false>(o, alpha[i], beta[i], M_ANY_ETA_POS, w.data(), &fl);
        ArrayW W{w.data()};
        bool reg = w[S_ETA] > 0.0;
        Imax_o[i] = reg ? w[S_TAUTOT] : -1.0;
        npol_o[i] = n_poloidal_at(W, w[S_TAUTOT]);
        for (int mbar = mlo; mbar <= mhi; ++mbar) {
            double rs = -1, Ir = -1, Nmax = -2;
            if (reg) {
                crossing_time(W, o, beta[i], mbar, &Ir, &Nmax);
                double d0, d1, d2;
                radial_eval<ArrayW, false>(W, o, fl, Ir, false, &rs, &d0, &d1, &d2);
                if (Nmax < (double)mbar) rs = 0.0;
            }
            size_t q = (size_t)(mbar - mlo) * n + i;
            rs_o[q] = rs; Ir_o[q] = Ir; Nmax_o[q] = Nmax;
        }
    }
}
void hs_counts(double a, double th_o, int n, int ntau, const double* alpha, const double* beta, const double* tau, double* neq, double* npol) {
    Observer o = make_observer(a, 1000.0, th_o);
    std::vector<double> w(NW);
    for (int i = 0; i < n; ++i) {
        int fl;
        ray_setup<true>(o, alpha[i], beta[i], M_ANY_ETA_POS, w.data(), &fl);
        ArrayW W{w.data()};
        for (int k = 0; k < ntau; ++k) {
            size_t q = (size_t)k * n + i;
            neq[q] = n_equatorial_at(W, o, fl, beta[i], tau[q]);
            npol[q] = n_poloidal_at(W, tau[q]);
        }
    }
}
void hs_velocity(const kgeo_models* M, double a, int n, const double* r, double* u) {
    for (int i = 0; i < n; ++i) { U4 v = velocity_u_lab(*M, a, r[i]); u[i] = v.u0; u[n + i] = v.u1; u[2 * n + i] = v.u2; u[3 * n + i] = v.u3; }
}
void hs_bfield(const kgeo_models* M, double a, int n, const double* r, double* B, double* bsq) {
    for (int i = 0; i < n; ++i) { bfield_lab(*M, a, r[i], B + i, B + n + i, B + 2 * n + i); bsq[i] = bfield_bsq(*M, a, r[i], velocity_u_lab(*M, a, r[i])); }
}
void hs_bfield_omega(const kgeo_models* M, double a, int n, const double* r, double* B, double* omega) {
    for (int i = 0; i < n; ++i) bfield_lab_omega(*M, a, r[i], B + i, B + n + i, B + 2 * n + i, omega + i);
}
void hs_emis(const kgeo_models* M, double a, int n, const double* r, const double* g, const double* s, const double* Bm, double nu, double specind, double* j) {
    for (int i = 0; i < n; ++i) j[i] = emissivity_jrest(*M, a, r[i], g != nullptr, g ? g[i] : 1.0, s != nullptr, s ? s[i] : 1.0, Bm ? Bm[i] : 0.0, nu, specind);
}
void hs_mino_total(double a, double r_o, int n, const double* roots, double* out) {
    Observer o = make_observer(a, r_o, 0.7);
    for (int i = 0; i < n; ++i) {
        Roots R;
        for (int k = 0; k < 4; ++k) { R.re[k] = roots[(size_t)k * n + i]; R.im[k] = roots[(size_t)(4 + k) * n + i]; }
        R.c12 = R.im[0] != 0.0; R.c34 = R.im[2] != 0.0; R.region = 0;
        out[i] = tau_total_of(o, R);
    }
}
}
