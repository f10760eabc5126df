/* kgeo_b200 -- C ABI of the B200-native analytic Kerr ray tracer (libkgeo_b200.so, sm_100a).
 *
 * This is the drop-in boundary for ONE hot path of achael/kgeo: the per-pixel analytic null-geodesic solve and the
 * equatorial image pipeline built on it.  The reference has no native code; its only FFI is ctypes -> libgsl for the
 * elliptic integral Pi (kgeo/gsl_ellip_binding.py:142-152, 3 doubles in -> 1 double out) and its "plugin seam" is the
 * module-level backend switch in kgeo/kerr_raytracing_ana.py:31-43.  Each entry point below names the reference
 * interface it replaces (file:line under /root/reference).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes.  Every function returns 0 on success, a positive cudaError_t value on a
 *    CUDA failure, or a negative KGEO_E_* code on bad arguments; kgeo_last_error() returns a static description.
 *  - `*_dev` pointers are DEVICE pointers owned by the caller (8-byte aligned); the library allocates nothing behind
 *    them.  The `*_host` entry points take HOST buffers, stage through device memory internally and return when the
 *    results are in the host buffers.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); device entry points are stream-ordered and
 *    asynchronous, thread-safe, and keep no global state.
 *  - Numerical trouble surfaces as NaN/inf in the outputs, never as an error code (as in the reference,
 *    kgeo/scipy_ellip_binding.py:138-162).
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails with cudaErrorNoDevice.
 */
#ifndef KGEO_B200_H
#define KGEO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGEO_B200_ABI_VERSION 4

#define KGEO_E_BADARG (-1)   /* NULL pointer, negative size, inconsistent shapes */
#define KGEO_E_RANGE (-2)    /* |a| >= 1, r_o < 100, th_o outside (0,pi) or == pi/2 (the reference's own checks) */
#define KGEO_E_MODEL (-3)    /* unknown physics model id */

/* Observer / black hole: (a, r_o, th_o) of observer_coords = [t, r_o, th_o, ph] (kgeo/kerr_raytracing_ana.py:46-71). */
typedef struct kgeo_observer {
    double a;     /* spin, |a| < 1 */
    double r_o;   /* observer radius >= 100, may be +inf for the lensing/image entry points */
    double th_o;  /* inclination in (0,pi), != pi/2 */
} kgeo_observer;

/* setup modes: the reference's whole-array guards np.any(eta>0), np.any(eta<=0) (kerr_raytracing_ana.py:217,300) and
 * the low-spin eta clamp that only the track entry points apply (:80-82) */
#define KGEO_MODE_ANY_ETA_POS 1
#define KGEO_MODE_ANY_ETA_NONPOS 2
#define KGEO_MODE_CLAMP_LOWSPIN_ETA 4
#define KGEO_MODE_INPUT_LAM_ETA 8 /* kgeo_ray_summary only: the two input arrays hold (lam, eta), not (alpha, beta) */
#define KGEO_MODE_POLISH_ROOTS 16 /* extension: Newton-polish the quartic roots (more accurate than the reference; off by default) */
#define KGEO_MODE_CARLSON_ONLY 32 /* evaluate E and Pi of every sample by Carlson duplication (the reference's algorithm); by default
                                     rays away from the critical curve use per-ray theta-function series (same values, ~2x faster) */

/* ---- physics models of the equatorial image (kgeo/velocities.py, kgeo/bfields.py, kgeo/emissivities.py) ---- */
enum kgeo_velocity_type { KGEO_VEL_ZAMO = 0, KGEO_VEL_INFALL, KGEO_VEL_KEP, KGEO_VEL_SUBKEP, KGEO_VEL_GENERAL,
                          KGEO_VEL_GELLES, KGEO_VEL_SIMFIT, KGEO_VEL_FROMFILE, KGEO_VEL_DRIFTFRAME };
enum kgeo_bfield_type { KGEO_B_SIMPLE = 0, KGEO_B_SIMPLE_RM1, KGEO_B_BZ_GUESS, KGEO_B_GEN_POWER, KGEO_B_MONOPOLE_A,
                        KGEO_B_FROMFILE, KGEO_B_BZ_MONOPOLE, KGEO_B_BZ_PARA, KGEO_B_POWER };
enum kgeo_emissivity_type { KGEO_EM_CONSTANT = 0, KGEO_EM_BPL, KGEO_EM_GLM, KGEO_EM_RING, KGEO_EM_THERMAL };

/* Parameter meaning per type (defaults as in the reference constructors):
 *  velocity: ZAMO/INFALL: none.  KEP: p[0]=retrograde.  SUBKEP: p[0]=retrograde, p[1]=fac_subkep.
 *            GENERAL: p[0]=retrograde, p[1]=fac_subkep, p[2]=beta_phi, p[3]=beta_r.  GELLES: p[0]=beta, p[1]=chi.
 *            SIMFIT: p[0]=ell_isco, p[1]=vr_isco, p[2]=p1, p[3]=p2, p[4]=dd.                 (velocities.py:26-68)
 *            DRIFTFRAME: p[0]=nu_parallel (numeric, no gammamax); the E x B drift of THIS struct's bfield model, which
 *            must be one with a field-line angular speed (velocities.py:487-623).
 *  bfield:   SIMPLE/SIMPLE_RM1: p[0]=Cr, p[1]=Cvert, p[2]=Cph ('rad','vert','tor' are SIMPLE with unit coefficients).
 *            BZ_GUESS: p[0]=C.  GEN_POWER: p[0]=n_I, p[1]=p_val, p[2]=isAbove.  MONOPOLE_A: p[0]=C, p[1]=omegafac.
 *            BZ_MONOPOLE: p[0]=C, p[1]=secondorder_only; bfield_table_dev = 2 rows x table_n doubles (r, f(r)), the
 *            reference's bz_fr_data.dat, evaluated as a linear spline with linear extrapolation (bfields.py:23-30,545-592).
 *            BZ_PARA: p[0]=C (Lambert-W field-line speed, bfields.py:609-673).  POWER: p[0]=p, p[1]=C, p[2]=usemono
 *            (bfields.py:618-625,676-709).
 *                                                                                            (bfields.py:36-91)
 *  emissivity: BPL: p[0]=p1, p[1]=p2.  GLM: p[0]=gamma_off, p[1]=sigma.  RING: p[0]=r_ring, p[1]=gamma_off, p[2]=sigma.
 *            THERMAL: p[0]=Rb, p[1]=ne0, p[2]=Te0, p[3]=B0, p[4]=alpha_n, p[5]=alpha_T, p[6]=alpha_B, p[7]=bfield_model.
 *            emiscut_in / emiscut_out as in Emissivity.__init__.                              (emissivities.py:39-72)
 */
typedef struct kgeo_models {
    int32_t velocity_type;
    int32_t bfield_type;
    int32_t emissivity_type;
    int32_t reserved;
    double velocity_p[8];
    double bfield_p[8];
    double emissivity_p[8];
    double emiscut_in;
    double emiscut_out;
    /* 'fromfile' models (kgeo/velocities.py:626-698, kgeo/bfields.py:749-806): DEVICE tables of 4 rows x table_n doubles,
     * row 0 = radii (increasing), rows 1..3 = U1,U2,U3 resp. Br,Btheta,Bphi, evaluated with numpy.interp semantics
     * (linear, clamped to the end values).  NULL / 0 for every other model. */
    const double* velocity_table_dev;
    const double* bfield_table_dev;
    int64_t velocity_table_n;
    int64_t bfield_table_n;
} kgeo_models;

/* Image options = keyword arguments of Iobs / make_image (kgeo/equatorial_images.py:23-32,124-127). */
typedef struct kgeo_image_opts {
    int32_t mbar_min;      /* orders mbar_min..mbar_max are produced (make_image: 0..mbar_max; Iobs: mbar..mbar) */
    int32_t mbar_max;
    int32_t polarization;  /* 0/1 */
    int32_t pathlength;    /* 0/1 */
    double specind;
    double diskangle;
    double nu_obs;
} kgeo_image_opts;

/* Where the pixels of a row shard go (kgeo_equatorial_image_sharded).  The shard is npix / cols whole image rows; local row lr is
 * written to row lr * row_step + row_offset of every destination plane (rows interleaved over N ranks: row_step = N,
 * row_offset = rank; a contiguous block of rows: row_step = 1, row_offset = first row; a whole image: cols = npix, 1, 0).
 * Consecutive order planes of one output are plane_stride elements apart in the destination (>= the destination plane). */
typedef struct kgeo_shard_map {
    uint64_t cols;
    uint64_t row_step;
    uint64_t row_offset;
    uint64_t plane_stride;
    uint64_t reserved;     /* 0 */
} kgeo_shard_map;

/* Exported CUDA IPC handle of a peer buffer (cudaIpcMemHandle_t, 64 bytes): plain data, sent to the other processes of the
 * box by whatever transport the host program has. */
typedef struct kgeo_peer_handle {
    unsigned char bytes[64];
} kgeo_peer_handle;

int kgeo_abi_version(void);
const char* kgeo_last_error(void);
/* number of CUDA devices visible to the library, or a negative/positive error code */
int kgeo_device_count(int* count);
/* Ray-constant workspace.  kgeo_ray_workspace_slots(): doubles per ray.  Layout: rays in groups of four (four consecutive
 * pixels = one 32-byte sector of every output row), group g = the contiguous block w[g*slots*4 ...) laid out [slot][ray in
 * group], i.e. slot s of ray i at w[(i/4)*slots*4 + s*4 + i%4].  A buffer for npix rays holds kgeo_ray_workspace_doubles(npix)
 * doubles (npix rounded up to a multiple of four rays) and must be 16-byte aligned (the track pass fetches a group's block
 * with one cp.async.bulk).  kgeo_ray_slot copies one slot out as a plain [npix] array (slot 2 = total Mino time). */
int kgeo_ray_workspace_slots(void);
size_t kgeo_ray_workspace_doubles(size_t npix);
int kgeo_ray_slot(size_t npix, const double* w_dev, int slot, double* out_dev, void* stream);
/* cumulative number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
long long kgeo_launch_count(void);

/* ---- element-wise special functions (device pointers) --------------------------------------------------------
 * Replace scipy.special.elliprf/elliprd/elliprj/elliprc, ellipkinc/ellipeinc/ellipk, ellipj as used at the call sites
 * listed in SURVEY.md 8(c), the libgsl bindings gsl_sf_ellint_RF_e/RJ_e/RC_e (kgeo/gsl_ellip_binding.py:142-152), and
 * ellip_pi_arr(n,phi,m) (kgeo/scipy_ellip_binding.py:87-237; same special-value conventions). */
int kgeo_special_rf(size_t n, const double* x_dev, const double* y_dev, const double* z_dev, double* out_dev, void* stream);
int kgeo_special_rd(size_t n, const double* x_dev, const double* y_dev, const double* z_dev, double* out_dev, void* stream);
int kgeo_special_rj(size_t n, const double* x_dev, const double* y_dev, const double* z_dev, const double* p_dev,
                    double* out_dev, void* stream);
int kgeo_special_rc(size_t n, const double* x_dev, const double* y_dev, double* out_dev, void* stream);
int kgeo_special_ellipk(size_t n, const double* m_dev, double* out_dev, void* stream);
int kgeo_special_ellipkinc(size_t n, const double* phi_dev, const double* m_dev, double* out_dev, void* stream);
int kgeo_special_ellipeinc(size_t n, const double* phi_dev, const double* m_dev, double* out_dev, void* stream);
int kgeo_special_ellippi(size_t n, const double* n_dev, const double* phi_dev, const double* m_dev, double* out_dev,
                         void* stream);
int kgeo_special_ellipj(size_t n, const double* u_dev, const double* m_dev, double* sn_dev, double* cn_dev,
                        double* dn_dev, double* am_dev, void* stream);

/* ---- per-ray setup --------------------------------------------------------------------------------------------
 * Replaces radial_roots, mino_total, uplus_uminus, angular_turning (kgeo/kerr_raytracing_utils.py:242-548) and the
 * observer-side antiderivatives of th_integrate / r_integrate (kgeo/kerr_raytracing_ana.py:231-251,312-327,431-434,
 * 508-510,583-584).  w_dev: kgeo_ray_workspace_doubles(npix) doubles; flags_dev: npix int32. */
int kgeo_ray_setup(const kgeo_observer* obs, size_t npix, const double* alpha_dev, const double* beta_dev, int mode,
                   double* w_dev, int32_t* flags_dev, void* stream);
/* The same setup from the conserved quantities themselves: lam, eta and s_o = the sign of the observer's polar momentum
 * (NULL: +1 everywhere), as th_integrate / r_integrate receive them (kgeo/kerr_raytracing_ana.py:195-196,368-369).
 * phi_and_t = 0 skips the constants only G_phi, G_t, I_phi, I_t, I_sigma need (do_phi_and_t=False; required for r_o = inf). */
int kgeo_ray_setup_conserved(const kgeo_observer* obs, size_t npix, const double* lam_dev, const double* eta_dev,
                             const double* s_o_dev, int mode, int phi_and_t, double* w_dev, int32_t* flags_dev, void* stream);
/* th_integrate (:195-365) and r_integrate (:368-633) as separate results at user Mino times tau_dev [ntau, tau_ld]:
 * (th, G_phi, G_t) and (r, I_phi, I_t, I_sigma), each [ntau, ld]; any output may be NULL, a whole side (polar / radial) whose
 * outputs are all NULL is not evaluated.  phi_and_t = 0: only th and r are computed; the others are zero except I_t = -4 tau (as the reference's). */
int kgeo_track_integrals(const kgeo_observer* obs, size_t npix, const double* w_dev, const int32_t* flags_dev, size_t ntau,
                         const double* tau_dev, size_t tau_ld, int phi_and_t, double* th_dev, double* Gph_dev, double* Gt_dev,
                         double* r_dev, double* Iph_dev, double* It_dev, double* Isig_dev, size_t ld, void* stream);
/* Unpacked per-ray summary for the host API: roots_dev[8*npix] = Re r1..r4, Im r1..r4 (plane-major), rclass_dev = the
 * reference's region label 1..4 (0 on a boundary), tau_tot_dev, u_plus/u_minus/uratio/a2u_minus, outside_dev =
 * is_outside_crit, thplus/thminus = angular_turning's turning points.  Any output may be NULL.  Unlike the geodesic entry points
 * this one takes any inclination 0 <= th_o <= pi (the reference's uplus_uminus / angular_turning do not restrict it).
 * (kerr_raytracing_utils.py:242-328,330-380,382-548,794-826) */
int kgeo_ray_summary(const kgeo_observer* obs, size_t npix, const double* alpha_dev, const double* beta_dev, int mode,
                     double* roots_dev, int32_t* rclass_dev, double* tau_tot_dev, double* uplus_dev, double* uminus_dev,
                     double* uratio_dev, double* a2uminus_dev, double* outside_dev, double* thplus_dev,
                     double* thminus_dev, void* stream);
/* mino_total(a, r_o, eta, r1..r4) from caller-supplied roots (kerr_raytracing_utils.py:382-548):
 * roots_dev[8*n] = Re r1..r4, Im r1..r4 plane-major; out_dev[n]. */
int kgeo_mino_total(const kgeo_observer* obs, size_t n, const double* roots_dev, double* out_dev, void* stream);

/* ---- geodesic tracks --------------------------------------------------------------------------------------------
 * Replaces coords_at_tau / raytrace_ana / th_integrate / r_integrate (kgeo/kerr_raytracing_ana.py:46-193,195-633).
 * One thread per (ray, Mino-time sample).  Outputs are [ntau, ld] row-major (sample-major, ray fastest), exactly the
 * reference's (ngeo, npix) arrays when ld == npix; any output pointer may be NULL (not materialised).
 *  tau_dev == NULL : samples are numpy.linspace(0, maxtaufrac*tau_tot, ntau) per ray (raytrace_ana :93-94);
 *  tau_dev != NULL : [ntau, tau_ld] user Mino times (coords_at_tau).
 *  full == 0       : only r and theta are computed (do_phi_and_t=False), t/phi/sigma outputs are zero-filled;
 *  full == 1       : all coordinates; with tau_dev == NULL two launches (theta-function-only hot pass over the four-ray groups
 *                    whose rays all qualify, complete code for the other groups);
 *  full == 2       : all coordinates, every ray through the complete per-point code in one launch (A/B reference: r and theta
 *                    are bit-identical to full == 1, phi / t / sigma agree to the last bits). */
int kgeo_tracks(const kgeo_observer* obs, size_t npix, const double* w_dev, const int32_t* flags_dev, size_t ntau,
                double maxtaufrac, const double* tau_dev, size_t tau_ld, int full, double* t_dev, double* r_dev,
                double* th_dev, double* ph_dev, double* sig_dev, double* tau_out_dev, size_t ld, void* stream);
/* Fused convenience: setup + tracks in one call, device pointers; w/flags scratch supplied by the caller. */
int kgeo_raytrace(const kgeo_observer* obs, size_t npix, const double* alpha_dev, const double* beta_dev, int mode,
                  size_t ngeo, double maxtaufrac, double* w_dev, int32_t* flags_dev, double* t_dev, double* r_dev,
                  double* th_dev, double* ph_dev, double* sig_dev, double* tau_out_dev, size_t ld, void* stream);
/* Host-buffer entry point (the end-to-end path): alpha/beta and all outputs are HOST pointers, [ngeo, npix]. */
int kgeo_raytrace_host(const kgeo_observer* obs, size_t npix, const double* alpha, const double* beta, int mode,
                       size_t ngeo, double maxtaufrac, double* t, double* r, double* th, double* ph, double* sig,
                       double* tau_out, int device);

/* ---- crossing counts (kgeo/kerr_raytracing_utils.py:551-684) ---------------------------------------------------
 * n_equatorial_crossings / n_poloidal_orbits at Mino times tau_dev[ntau, npix]; outputs [ntau, npix], may be NULL. */
int kgeo_crossing_counts(const kgeo_observer* obs, size_t npix, const double* beta_dev, const double* w_dev,
                         const int32_t* flags_dev, size_t ntau, const double* tau_dev, double* neq_dev,
                         double* npol_dev, void* stream);

/* ---- equatorial lensing (kgeo/equatorial_lensing.py:22-196) ----------------------------------------------------
 * r_equatorial for mbar = mbar_lo .. mbar_hi in one pass: outputs are [(mbar_hi-mbar_lo+1), npix] planes
 * (r_s, Ir, Nmax per order) and [npix] (Imax, n_poloidal = nmax_poloidal); any may be NULL. */
int kgeo_r_equatorial(const kgeo_observer* obs, size_t npix, const double* alpha_dev, const double* beta_dev,
                      int mbar_lo, int mbar_hi, double* rs_dev, double* Ir_dev, double* Nmax_dev, double* Imax_dev,
                      double* npol_dev, void* stream);

/* ---- inverse lensing (kgeo/equatorial_lensing.py:197-298) -----------------------------------------------------------
 * rho_of_req: for every image-plane polar angle varphi_dev[i] the radius rho at which the mbar-th equatorial crossing of the
 * ray (alpha, beta) = rho (cos varphi, sin varphi) happens at r = req, observer at infinity (the reference's R0).  One thread
 * per angle runs the bracketed root find the reference runs serially: SciPy's brentq (scipy/optimize/Zeros/brentq.c, SciPy
 * 1.18: inverse quadratic interpolation / secant / bisection, tolerance (xtol + rtol |x|)/2) on r_equatorial2(rho) - req with
 * the reference's sign-definite patches (r <= 0 -> 1e50 outside the critical curve, 1 inside; :208-215).
 * status_dev[i]: 0 converged, -1 f(rho_min) and f(rho_max) have the same sign (SciPy raises ValueError), -2 not converged
 * after maxiter iterations (SciPy raises RuntimeError).  rho/alpha/beta outputs may be NULL.  r_o of obs is ignored. */
int kgeo_rho_of_req(const kgeo_observer* obs, size_t n, const double* varphi_dev, double req, int mbar, double rho_min,
                    double rho_max, double xtol, double rtol, int maxiter, double* rho_dev, double* alpha_dev,
                    double* beta_dev, int32_t* status_dev, void* stream);

/* A library of such curves in one launch: generate_library / make_curves2 (:351-417; the reference's only multi-process job:
 * a pool over 99 x 89 (spin, inclination) configurations, each running rho_of_req for mbar = 0..5 at 360 angles).  Configuration
 * c has spin spin_dev[c], inclination th0_dev[c] (radians) and ring radius req_dev[c]; orders mbar_lo..mbar_hi with the bracket
 * rho_of_req picks for each (:232-241).  rho_dev / status_dev: [nconf, norder, nphi]; status as kgeo_rho_of_req (the caller
 * decides what a failed root find means: the reference drops the configuration's file).  The spins must satisfy |a| < 1 and
 * the inclinations 0 < th0 < pi, th0 != pi/2 (checked by the caller: the arrays are device memory). */
int kgeo_rho_library(size_t nconf, const double* spin_dev, const double* th0_dev, const double* req_dev, int mbar_lo, int mbar_hi,
                     size_t nphi, const double* varphi_dev, double xtol, double rtol, int maxiter, double* rho_dev,
                     int32_t* status_dev, void* stream);

/* critical_curve (:268-298, GL19 eqs 38-40): (alpha, beta) of the n photon-orbit radii r = linspace(r_minus, r_plus, n);
 * the caller passes the scalars the reference computes with NumPy (r_minus, r_plus, sin/cos/tan th_0) so that the NaN mask
 * next to the end points is the reference's.  Outputs [n]; NaN entries are dropped and the curve mirrored by the caller. */
int kgeo_critical_curve(double a, double r_minus, double r_plus, double sin_th0, double cos_th0, double tan_th0, size_t n,
                        double* alpha_dev, double* beta_dev, void* stream);

/* ---- equatorial image (kgeo/equatorial_images.py:23-544) -------------------------------------------------------
 * n = mbar_min..mbar_max image in one call (two launches: geometry, then emission): per pixel one setup, then per order the crossing radius, redshift, (polarization,)
 * emissivity and Stokes I,Q,U.  Planes are [mbar_max-mbar_min+1, npix] (order-major); the Python wrapper exposes them with the
 * reference's (npix, mbar_max+1) indexing.  nmax_dev = Nmax of the last order, npol_dev = nmax_poloidal, imax_dev = the
 * total Mino time (Iobs's Imax; -1 for vortical rays); lp/Ie only when opts->pathlength.  Any output may be NULL. */
int kgeo_equatorial_image(const kgeo_observer* obs, const kgeo_models* models, const kgeo_image_opts* opts, size_t npix,
                          const double* alpha_dev, const double* beta_dev, double* I_dev, double* Q_dev, double* U_dev,
                          double* r_dev, double* Ir_dev, double* g_dev, double* sinthb_dev, double* nmax_dev,
                          double* npol_dev, double* imax_dev, double* lp_dev, double* Ie_dev, void* stream);
/* Row-sharded image with the gather fused into the kernel (BASELINE cfg 4 / 5 over the GPUs of a box; the reference computes
 * the whole plane in one process, kgeo/equatorial_images.py:23-122): as kgeo_equatorial_image for the npix pixels of this
 * shard, but every output element is stored at its place in destination planes described by `map` -- which may be memory of
 * ANOTHER GPU of the box mapped with kgeo_peer_open (coalesced stores over NVLink; no staging copy, no separate collective,
 * no un-interleave pass).  The [npix] outputs (nmax, npol, imax) use the same row map.  The stores are complete when the
 * kernel is; the owner of the destination must still order its reads after the writers' streams (any cross-rank barrier). */
int kgeo_equatorial_image_sharded(const kgeo_observer* obs, const kgeo_models* models, const kgeo_image_opts* opts, size_t npix,
                                  const double* alpha_dev, const double* beta_dev, const kgeo_shard_map* map, double* I_dst,
                                  double* Q_dst, double* U_dst, double* r_dst, double* Ir_dst, double* g_dst,
                                  double* sinthb_dst, double* nmax_dst, double* npol_dst, double* imax_dst, double* lp_dst,
                                  double* Ie_dst, void* stream);
/* The same placement for planes that were computed locally first: src_dev [nplanes, npix] (plane-major, this shard's rows) is
 * copied to its rows of the destination planes by the COPY ENGINES (one strided 3-D cudaMemcpy; destination may be peer memory),
 * so the transfer runs beside kernels of another stream without taking SMs from them. */
int kgeo_shard_scatter(const double* src_dev, size_t nplanes, size_t npix, const kgeo_shard_map* map, double* dst, void* stream);
/* Peer buffers.  kgeo_peer_alloc: `bytes` of device memory on the current device plus the handle under which the other
 * processes of the box open it; kgeo_peer_open (in another process, on ITS current device): the address at which that
 * device reads and writes the exporter's memory (peer access over NVLink is enabled on first use); kgeo_peer_close unmaps,
 * kgeo_peer_free releases the allocation in the exporting process (after every importer has closed it). */
int kgeo_peer_alloc(size_t bytes, void** dev_ptr, kgeo_peer_handle* handle);
int kgeo_peer_open(const kgeo_peer_handle* handle, void** dev_ptr);
int kgeo_peer_close(void* dev_ptr);
int kgeo_peer_free(void* dev_ptr);
/* Host-buffer variant (end-to-end path).  Same layouts, HOST pointers. */
int kgeo_equatorial_image_host(const kgeo_observer* obs, const kgeo_models* models, const kgeo_image_opts* opts,
                               size_t npix, const double* alpha, const double* beta, double* I, double* Q, double* U,
                               double* r, double* Ir, double* g, double* sinthb, double* nmax, double* npol,
                               double* imax, double* lp, double* Ie, int device);

/* ---- element-wise physics closed forms (device pointers) -------------------------------------------------------
 * Velocity.u_lab (kgeo/velocities.py:70-96), Bfield.bfield_lab / Bfield.bsq (kgeo/bfields.py:93-131,365-396),
 * Emissivity.jrest (kgeo/emissivities.py:83-118) in the equatorial plane.  g/sinthb/Bmag may be NULL in jrest. */
int kgeo_velocity_u_lab(const kgeo_models* models, double a, size_t n, const double* r_dev, double* u0_dev,
                        double* u1_dev, double* u2_dev, double* u3_dev, void* stream);
int kgeo_bfield_lab(const kgeo_models* models, double a, size_t n, const double* r_dev, double* B1_dev, double* B2_dev,
                    double* B3_dev, double* bsq_dev, void* stream);
/* Bfield.omega_field (kgeo/bfields.py:175-197): field-line angular speed of the models that define one (bz_monopole,
 * monopoleA, bz_guess, bz_para, power, gen_power); NaN for the others. */
int kgeo_bfield_omega(const kgeo_models* models, double a, size_t n, const double* r_dev, double* omega_dev, void* stream);
int kgeo_emissivity_jrest(const kgeo_models* models, double a, size_t n, const double* r_dev, const double* g_dev,
                          const double* sinthb_dev, const double* Bmag_dev, double nu_obs, double specind,
                          double* j_dev, void* stream);

/* The host-buffer entry points take their device scratch from the device's default memory pool and leave it cached there
 * between calls (a cfg 3 trace uses ~4 GB; returning it to the driver cost up to 200 ms per call).  kgeo_trim_memory gives
 * the cached blocks back (synchronises the device). */
int kgeo_trim_memory(int device);

/* ---- field-line crossings of the jet models (kgeo/solver.py:98-207: getguesses / findroot) --------------------------
 * psi models: 0 = BZ paraboloid, 1 = monopole, 2 = power law r^pval (1 - |cos theta|), 3 = none (psi = 0).
 * kgeo_fieldline_guesses, maxlen <= 0: per ray the number of starting guesses the reference's sign-change scan over the
 *   (ngeo, npix) track arrays r, theta, tau yields (>= 1: a ray without crossing holds the single guess -1) -> count_dev,
 *   their maximum -> maxcount_dev, the search limit of get_maxtau_forwardjet -> taumax_dev (optional); neqmax < 0 = None.
 *   ngeo = 0 (track pointers may be NULL) skips the scan: only taumax_dev is of interest then.
 *   maxlen > 0: the guesses themselves, guesses_dev[m * npix + ray], padded with -1 (makegoodarray + transpose).
 * kgeo_fieldline_solve: scipy.optimize.newton's vectorised secant iteration on psi(r(tau), theta(tau)) - psitarget for the
 *   nguess x npix guesses (all elements iterate until every one has converged, as in the reference), then r, theta at the
 *   roots, the signs of p^r and p^theta and the crossing count of getneq; elements whose guess is -1 come back as zeros.
 *   Blocks until done (the iteration count depends on device-side convergence flags).  iterations / nfailed: host ints.
 *   nfailed (only when maxiter ran out): the number of elements scipy.optimize.newton reports as failed (|dp| >= tol at the
 *   last step, or never updated), NEGATED when some of them had a zero secant derivative at the last step (SciPy then only
 *   warns; with all elements failed and none of them zero-derivative it raises RuntimeError). */
int kgeo_fieldline_guesses(const kgeo_observer* obs, size_t npix, size_t ngeo, const double* r_dev, const double* th_dev,
                           const double* tau_dev, size_t ld, const double* alpha_dev, const double* beta_dev, const double* w,
                           const int32_t* flags, int model, double pval, double psitarget, int neqmax, double maxtaufrac,
                           int maxlen, int32_t* count_dev, int32_t* maxcount_dev, double* guesses_dev, double* taumax_dev,
                           void* stream);
int kgeo_fieldline_solve(const kgeo_observer* obs, size_t npix, size_t nguess, const double* beta_dev, const double* w,
                         const int32_t* flags, int model, double pval, double psitarget, const double* guesses_dev, double tol,
                         int maxiter, double* tau_dev, double* r_dev, double* th_dev, double* signpr_dev, double* signpth_dev,
                         double* neq_dev, int32_t* iterations, int32_t* nfailed, void* stream);

/* ---- off-equatorial jet emission (kgeo/off_eq.py:124-233 getstokes, kgeo/densityfuncs.py, kgeo/geometry.py:38-62) -------------
 * kgeo_jet_stokes: for the nguess x npix crossings of kgeo_fieldline_solve (r, theta, sign p^r, sign p^theta; r = 0 = none) the
 *   density of the jet model (field 0 = BZ monopole with its f(r) table of 2 x fr_n doubles, 1 = power law r^pval; r0 =
 *   stagnation radius, norm = source integral from r0 to infinity for the power law) and Iobs_off's polarized Stokes parameters
 *   with Emissivity('constant') in the drift frame of the field; NaN -> 0 as in getstokes.
 *   Parallel boost of the drift frame (kgeo/velocities.py:474-516): nu_ff == 0: the number nu_parallel; nu_ff != 0:
 *   nu_parallel = 'FF', the energy-conserving boost of getnu_cons, with the field-line constants (r0, omega, E - Omega_F L =
 *   ff_boost.getEco) of Iobs_off's velocity in ff_global[3] (host) and those of the density's velocity, one triple per crossing
 *   order, in ff_rows_dev[3 * nguess] (device).  gammamax > 0: Lorentz factor capped as in :593-611 (gamma_eff rescaled at the
 *   crossing of smallest gamma_eff: a device-side argmin pre-pass).  *nu_out_of_range (host int, optional) is set when a boost
 *   left [-1, 1], where the reference raises.  Blocks until done.
 * kgeo_jet_sort: geometry.sort_image: (neqmax + 1) x npix planes, plane k < neqmax = crossings of order k, last plane = all. */
int kgeo_jet_stokes(const kgeo_observer* obs, size_t npix, size_t nguess, const double* alpha_dev, const double* beta_dev,
                    const double* r_dev, const double* th_dev, const double* signpr_dev, const double* signpth_dev, int field, double C,
                    double pval, int usemono, const double* fr_table_dev, int64_t fr_n, double nu_parallel, int nu_ff,
                    const double* ff_global, const double* ff_rows_dev, double gammamax, double sigma, double r0, double norm,
                    double psitarget, double specind, double* I_dev, double* Q_dev, double* U_dev, double* sinthb_dev,
                    double* density_dev, int32_t* nu_out_of_range, void* stream);
int kgeo_jet_sort(size_t npix, size_t nguess, int neqmax, const double* I_dev, const double* Q_dev, const double* U_dev,
                  const double* neq_dev, double* out_I_dev, double* out_Q_dev, double* out_U_dev, void* stream);

/* ---- measurement helper: sustained DFMA rate of the device (the measured FP64 roofline denominator) ------------ */
int kgeo_fp64_peak_probe(int iters, double* flops_per_launch, float* ms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KGEO_B200_H */
