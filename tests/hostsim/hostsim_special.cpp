// DEVELOPER-ONLY harness: compiles the device math headers with the HOST compiler so the algorithms can be
// checked against the oracle in a container without a GPU.  Never loaded by kgeo_b200; not a CPU fallback.
#include "../../kgeo_b200/csrc/kg_special.cuh"
extern "C" {
void hs_rf(int n, const double* x, const double* y, const double* z, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::carlson_rf(x[i], y[i], z[i]); }
void hs_rd(int n, const double* x, const double* y, const double* z, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::carlson_rd(x[i], y[i], z[i]); }
void hs_rj(int n, const double* x, const double* y, const double* z, const double* p, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::carlson_rj(x[i], y[i], z[i], p[i]); }
void hs_rc(int n, const double* x, const double* y, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::carlson_rc(x[i], y[i]); }
void hs_F(int n, const double* phi, const double* m, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::ellip_F(phi[i], m[i]); }
void hs_E(int n, const double* phi, const double* m, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::ellip_E(phi[i], m[i]); }
void hs_K(int n, const double* m, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::ellip_K(m[i]); }
void hs_Pi(int n, const double* nn, const double* phi, const double* m, double* o) { for (int i = 0; i < n; ++i) o[i] = kg::ellip_Pi_ref(nn[i], phi[i], m[i]); }
void hs_jacobi(int n, const double* u, const double* m, double* sn, double* cn, double* dn, double* am) {
    for (int i = 0; i < n; ++i) { double K; kg::jacobi_sncndn(u[i], m[i], sn + i, cn + i, dn + i, &K); am[i] = kg::jacobi_am(u[i], sn[i], cn[i], K); }
}
}
