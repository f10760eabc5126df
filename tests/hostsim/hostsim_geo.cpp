// DEVELOPER-ONLY harness (see hostsim_special.cpp): host compile of the device geodesic math for algorithm checks.
#include "../../kgeo_b200/csrc/kg_geodesic.cuh"
#include <vector>
extern "C" {
int hs_nw() { return kg::NW; }
void hs_setup(double a, double r_o, double th_o, int mode, int n, const double* alpha, const double* beta, double* w, int* flags) {
    kg::Observer o = kg::make_observer(a, r_o, th_o);
    for (int i = 0; i < n; ++i) kg::ray_setup<true>(o, alpha[i], beta[i], mode, w + (size_t)i * kg::NW, flags + i);
}
// tau: [ntau, n]; outputs [ntau, n]
void hs_points(double a, double r_o, double th_o, int mode, int full, int n, int ntau, const double* alpha, const double* beta,
               const double* tau, double* t, double* r, double* th, double* ph, double* sig) {
    kg::Observer o = kg::make_observer(a, r_o, th_o);
    std::vector<double> w(kg::NW);
    for (int i = 0; i < n; ++i) {
        int fl;
        kg::ray_setup<true>(o, alpha[i], beta[i], mode, w.data(), &fl);
        kg::ArrayW W{w.data()};
        for (int k = 0; k < ntau; ++k) {
            kg::Point p = kg::point_eval(W, o, fl, tau[(size_t)k * n + i], full != 0);
            size_t q = (size_t)k * n + i;
            t[q] = p.t; r[q] = p.r; th[q] = p.th; ph[q] = p.ph; sig[q] = p.sig;
        }
    }
}
}
