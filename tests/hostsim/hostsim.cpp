// Developer-only host harness: compiles the __host__ __device__ per-ray / per-sample functions of kgeo_b200/csrc with the
// HOST compiler so that algorithm changes can be compared with the oracle without a GPU.  It is NOT part of the product:
// libkgeo_b200.so never contains or calls this, and kgeo_b200/ never loads it (the product has no CPU path).
//   g++ -O2 -ffp-contract=off -shared -fPIC -I../../kgeo_b200/csrc hostsim.cpp -o _hostsim.so
#include <cstddef>
#include <cstdint>
#include "kg_geodesic.cuh"
#include "kg_hot.cuh"

using namespace kg;

extern "C" {

int hostsim_slots() { return NW; }
// element-wise access to the hot-loop elementary functions of kg_math.cuh (which: 0 log, 1 sin, 2 cos, 3 atan2(x, y), 4 acos, 5 atanh_ratio(x, y))
void hostsim_math(int which, size_t n, const double* x, const double* y, double* out) {
    for (size_t i = 0; i < n; ++i) {
        double s, c;
        switch (which) {
            case 0: out[i] = flog(x[i]); break;
            case 1: fsincos(x[i], &s, &c); out[i] = s; break;
            case 2: fsincos(x[i], &s, &c); out[i] = c; break;
            case 3: out[i] = fatan2(x[i], y[i]); break;
            case 5: out[i] = atanh_ratio(x[i], y[i]); break;
            default: out[i] = facos(x[i]); break;
        }
    }
}
// the setup as the track kernels launch it (DEFER = true): whole (PART 0) and as its two parts (PART 1 then PART 2) into
// workspaces pre-filled with a sentinel; [npix][NW] each.  Bit-identical by construction; checked by tests/test_abi_and_host.py.
int hostsim_setup_parts(double a, double r_o, double th_o, size_t npix, const double* alpha, const double* beta, int mode,
                        double* w_whole, int32_t* fl_whole, double* w_parts, int32_t* fl_parts) {
    Observer o = make_observer(a, r_o, th_o);
    for (size_t i = 0; i < npix; ++i) {
        int f0 = 0, f1 = 0;
        ray_setup<true, true, true, double*, 0>(o, alpha[i], beta[i], mode, w_whole + i * NW, &f0);
        ray_setup<true, true, true, double*, 1>(o, alpha[i], beta[i], mode, w_parts + i * NW, &f1);
        ray_setup<true, true, true, double*, 2>(o, alpha[i], beta[i], mode, w_parts + i * NW, &f1);
        fl_whole[i] = f0;
        fl_parts[i] = f1;
    }
    return 0;
}
int hostsim_slot(int which) { const int v[] = {T_TH, R_TH, R_K, T_M, S_TAUTOT, TB_SIZE, TP_SIZE, R_N0}; return v[which]; }

// tracks for npix rays, ntau samples each (numpy.linspace(0, maxtaufrac*tau_tot, ntau)); outputs [ntau][npix] row-major.
// force: 0 = flags as computed, 1 = clear the theta-function flags (Carlson path everywhere)
int hostsim_tracks(double a, double r_o, double th_o, size_t npix, const double* alpha, const double* beta, int mode,
                   int ntau, double maxtaufrac, int force, double* t, double* r, double* th, double* ph, double* sig,
                   int32_t* flags_out, double* w_out) {
    Observer o = make_observer(a, r_o, th_o);
    double w[NW];
    for (size_t i = 0; i < npix; ++i) {
        int fl;
        ray_setup<true, true>(o, alpha[i], beta[i], mode, w, &fl);
        if (force == 1) fl &= ~(F_TH_THETA | F_R_THETA);
        if (force == 2) fl &= ~F_TH_THETA;
        if (force == 3) fl &= ~F_R_THETA;
        if (flags_out) flags_out[i] = fl;
        if (w_out) for (int s = 0; s < NW; ++s) w_out[(size_t)s * npix + i] = w[s];
        ArrayW W{w};
        const double stop = maxtaufrac * w[S_TAUTOT];
        const double step = (ntau > 1) ? stop / (double)(ntau - 1) : 0.0;
        if (force == 4 && (fl & (F_TH_THETA | F_R_THETA)) == (F_TH_THETA | F_R_THETA)) {   // hot code, two samples at a time
            for (int k = 0; k < ntau; k += 2) {
                double tau[2];
                for (int s = 0; s < 2; ++s) {
                    int kk = (k + s < ntau) ? k + s : k;
                    tau[s] = (kk == ntau - 1 && ntau > 1) ? stop : (double)kk * step;
                }
                double th_[2], Gph[2], Gt[2], r_[2], Iph[2], It[2], Isig[2];
                polar_hot<ArrayW, 2>(W, fl, tau, th_, Gph, Gt);
                radial_hot<ArrayW, 2>(W, o, fl, tau, r_, Iph, It, Isig);
                for (int s = 0; s < 2 && k + s < ntau; ++s) {
                    size_t q = (size_t)(k + s) * npix + i;
                    t[q] = It[s] + o.a2 * Gt[s]; r[q] = r_[s]; th[q] = th_[s]; ph[q] = Iph[s] + w[S_LAM] * Gph[s];
                    sig[q] = Isig[s] + o.a2 * Gt[s];
                }
            }
            continue;
        }
        for (int k = 0; k < ntau; ++k) {
            double tau = (k == ntau - 1 && ntau > 1) ? stop : (double)k * step;
            Point p = point_eval(W, o, fl, tau, true);
            size_t q = (size_t)k * npix + i;
            t[q] = p.t; r[q] = p.r; th[q] = p.th; ph[q] = p.ph; sig[q] = p.sig;
        }
    }
    return 0;
}
}
