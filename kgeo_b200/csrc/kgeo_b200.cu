// libkgeo_b200.so -- CUDA kernels (sm_100a) and the C ABI declared in include/kgeo_b200.h.
//
// Kernels
//   ray_setup_kernel   one thread per ray: conserved quantities, quartic roots, root class, tau_tot, observer-side
//                      antiderivatives and complete integrals -> slot-major SoA workspace (coalesced in the ray index).
//   track_kernel       the hot kernel: one thread per (ray, Mino-time sample).  A CTA works through 8 groups of 4 consecutive
//                      rays; a group's constants are staged in shared memory ([slot][4]) and all warps of the CTA walk
//                      the group's samples together (lane = 4 * sample + ray: every store instruction writes 8 full
//                      32-byte sectors of the reference's (ngeo, npix) arrays).  A full trace is two launches: a hot
//                      pass without the Carlson fallback (64 registers) and a fix-up pass for the fallback rays.
//   r_equatorial_kernel + image_emission_kernel
//                      equatorial image in two launches: per pixel one setup in registers and the crossing time / radius of
//                      every order mbar_min..mbar_max (planes), then redshift, polarization, emissivity per order; no
//                      geodesic array ever goes through HBM.
//   plus element-wise kernels for the special functions and the physics closed forms.
//
// FP64 vector pipe bound; no tensor cores (there is no contraction in this path).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>

#include "../../include/kgeo_b200.h"
#include "kg_image.cuh"
#include "kg_hot.cuh"
#include "kg_fieldline.cuh"
#include "kg_offeq.cuh"

using namespace kg;

// ---------------------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail_arg(int code, const char* what) {
    snprintf(g_err, sizeof(g_err), "kgeo_b200: %s", what);
    return code;
}
static int fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "kgeo_b200: CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), where);
    return (int)e;
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);   \
    } while (0)
#define LAUNCHED(name)                                        \
    do {                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);   \
        cudaError_t e_ = cudaGetLastError();                  \
        if (e_ != cudaSuccess) return fail_cuda(e_, name);    \
    } while (0)

// any_inclination: the per-ray summary (uplus_uminus, angular_turning, radial_roots, mino_total: kerr_raytracing_utils.py:242-548)
// takes every th_o, as the reference's functions do (their inclination checks are commented out; n_angular_turnings calls them for
// an equatorial observer); the geodesic entry points need th_o off the equatorial plane (kerr_raytracing_ana.py:63-66)
static int check_observer(const kgeo_observer* obs, bool allow_inf, bool any_inclination = false) {
    if (!obs) return fail_arg(KGEO_E_BADARG, "observer is NULL");
    if (!(fabs(obs->a) < 1.0)) return fail_arg(KGEO_E_RANGE, "a should be a float in range (-1,1)");
    if (!(obs->r_o >= 100.0)) return fail_arg(KGEO_E_RANGE, "r_o should be a float > 100");
    if (!allow_inf && isinf(obs->r_o)) return fail_arg(KGEO_E_RANGE, "r_o must be finite for this entry point");
    if (any_inclination) {
        if (!(obs->th_o >= 0.0 && obs->th_o <= kPi)) return fail_arg(KGEO_E_RANGE, "th_o should be a float in range [0,pi]");
    } else if (!(obs->th_o > 0.0 && obs->th_o < kPi) || obs->th_o == 0.5 * kPi) {
        return fail_arg(KGEO_E_RANGE, "th_o should be a float in range (0,pi/2) or (pi/2,pi)");
    }
    return 0;
}
static int check_models(const kgeo_models* m) {
    if (!m) return fail_arg(KGEO_E_BADARG, "models is NULL");
    if (m->velocity_type < 0 || m->velocity_type > KGEO_VEL_DRIFTFRAME) return fail_arg(KGEO_E_MODEL, "unknown velocity model");
    if (m->bfield_type < 0 || m->bfield_type > KGEO_B_POWER) return fail_arg(KGEO_E_MODEL, "unknown bfield model");
    if (m->velocity_type == KGEO_VEL_FROMFILE && (!m->velocity_table_dev || m->velocity_table_n < 1))
        return fail_arg(KGEO_E_MODEL, "velocity 'fromfile' needs a device table (velocity_table_dev, velocity_table_n)");
    if (m->bfield_type == KGEO_B_BZ_MONOPOLE && (!m->bfield_table_dev || m->bfield_table_n < 2))
        return fail_arg(KGEO_E_MODEL, "bfield model 'bz_monopole' needs the device table of f(r)");
    if (m->bfield_type == KGEO_B_FROMFILE && (!m->bfield_table_dev || m->bfield_table_n < 1))
        return fail_arg(KGEO_E_MODEL, "bfield 'fromfile' needs a device table (bfield_table_dev, bfield_table_n)");
    if (m->emissivity_type < 0 || m->emissivity_type > KGEO_EM_THERMAL) return fail_arg(KGEO_E_MODEL, "unknown emissivity model");
    return 0;
}

static inline unsigned grid_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------------------------------------
// element-wise special-function kernels
// ---------------------------------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 128;

#define EW_INDEX                                                        \
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;           \
    if (i >= n) return;

__global__ void __launch_bounds__(EW_THREADS) ew_rf_kernel(size_t n, const double* x, const double* y, const double* z, double* o) {
    EW_INDEX
    o[i] = carlson_rf(x[i], y[i], z[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_rd_kernel(size_t n, const double* x, const double* y, const double* z, double* o) {
    EW_INDEX
    o[i] = carlson_rd(x[i], y[i], z[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_rj_kernel(size_t n, const double* x, const double* y, const double* z, const double* p, double* o) {
    EW_INDEX
    o[i] = carlson_rj(x[i], y[i], z[i], p[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_rc_kernel(size_t n, const double* x, const double* y, double* o) {
    EW_INDEX
    o[i] = carlson_rc(x[i], y[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_K_kernel(size_t n, const double* m, double* o) {
    EW_INDEX
    o[i] = ellip_K(m[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_F_kernel(size_t n, const double* phi, const double* m, double* o) {
    EW_INDEX
    o[i] = ellip_F(phi[i], m[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_E_kernel(size_t n, const double* phi, const double* m, double* o) {
    EW_INDEX
    o[i] = ellip_E(phi[i], m[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_Pi_kernel(size_t n, const double* nn, const double* phi, const double* m, double* o) {
    EW_INDEX
    o[i] = ellip_Pi_ref(nn[i], phi[i], m[i]);
}
__global__ void __launch_bounds__(EW_THREADS) ew_jacobi_kernel(size_t n, const double* u, const double* m, double* sn, double* cn,
                                                               double* dn, double* am) {
    EW_INDEX
    double s, c, d, K;
    jacobi_sncndn(u[i], m[i], &s, &c, &d, &K);
    if (sn) sn[i] = s;
    if (cn) cn[i] = c;
    if (dn) dn[i] = d;
    if (am) am[i] = jacobi_am(u[i], s, c, K);
}

// ---------------------------------------------------------------------------------------------------------------------
// per-ray setup
// ---------------------------------------------------------------------------------------------------------------------
#ifndef KG_SETUP_THREADS
#define KG_SETUP_THREADS 256
#endif
constexpr int SETUP_THREADS = KG_SETUP_THREADS;

#ifndef KG_SETUP_SPLIT
#define KG_SETUP_SPLIT 1              // the per-ray setup as two launches (polar | radial); 0: one
#endif
#ifndef KG_SETUP_MINBLOCKS
#define KG_SETUP_MINBLOCKS 2          // CTAs per SM the two setup kernels are compiled for (register cap = 65536/(threads*n))
#endif

// PART 1 = conserved quantities + polar motion, PART 2 = radial motion (two launches: the ~15.6 k instructions of the whole
// setup, walked once per warp, stalled on instruction fetch 3.4 cycles per issue; see ray_setup_conserved), 0 = both.
template <bool PHI_T, int PART>
__global__ void __launch_bounds__(SETUP_THREADS, KG_SETUP_MINBLOCKS) ray_setup_kernel(Observer o, size_t npix, const double* __restrict__ alpha,
                                                                  const double* __restrict__ beta, int mode,
                                                                  double* __restrict__ w_out, int32_t* __restrict__ flags_out) {
    // (Re-converging the CTA's warps with barriers at the stage boundaries of ray_setup, so that they share fetched instruction
    // lines, was measured: no change, 1.75 ms either way.)
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    int fl = (PART == 2) ? flags_out[i] : 0;
    ray_setup<PHI_T, true, true, WsPtr, PART>(o, alpha[i], beta[i], mode, ray_ptr(w_out, i), &fl);   // constants go straight to the grouped workspace
    flags_out[i] = fl;
}

// second setup kernel: theta-function blocks of the rays ray_setup_kernel left pending (see theta_setup)
__global__ void __launch_bounds__(SETUP_THREADS, KG_SETUP_MINBLOCKS) theta_setup_kernel(Observer o, size_t npix, double* __restrict__ w_io,
                                                                    int32_t* __restrict__ flags_io) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    int fl = flags_io[i];
    if (!(fl & (F_TH_PENDING | F_R_PENDING))) return;
    theta_setup(o, ray_ptr(w_io, i), &fl);
    flags_io[i] = fl;
}

// one slot of the grouped workspace as a plain [npix] array
__global__ void __launch_bounds__(SETUP_THREADS) ray_slot_kernel(size_t npix, const double* __restrict__ w, int slot, double* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) out[i] = ray_view(w, i)(slot);
}

__global__ void __launch_bounds__(SETUP_THREADS) ray_summary_kernel(Observer o, size_t npix, const double* __restrict__ alpha,
                                                                    const double* __restrict__ beta, int mode, double* roots,
                                                                    int32_t* rclass, double* tau_tot, double* uplus,
                                                                    double* uminus, double* uratio, double* a2uminus,
                                                                    double* outside, double* thplus, double* thminus) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    double al = alpha[i], be = beta[i];
    double lam, eta;
    if (mode & KGEO_MODE_INPUT_LAM_ETA) {   // the two input arrays are (lam, eta) instead of (alpha, beta)
        lam = al;
        eta = be;
    } else {
        lam = -al * o.sth;
        eta = conserved_eta(o, al, be);
    }
    if (o.lowspin && (mode & M_CLAMP_LOWSPIN_ETA) && eta < 0.0) eta = kEP;
    if (roots || rclass || outside) {
        Roots R = quartic_roots(o, lam, eta, (mode & M_POLISH_ROOTS) != 0);
        if (roots) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                roots[(size_t)k * npix + i] = R.re[k];
                roots[(size_t)(4 + k) * npix + i] = R.im[k];
            }
        }
        if (rclass) rclass[i] = R.region;
        if (outside) outside[i] = (!R.c34 && R.re[3] > o.rplus && !(eta < 0.0)) ? 1.0 : 0.0;
    }
    if (uplus || uminus || uratio || a2uminus || thplus || thminus) {
        double up, um, ratio, a2um;
        angular_consts(o, lam, eta, &up, &um, &ratio, &a2um);
        if (thplus) thplus[i] = (eta < 0.0) ? acos(sqrt(um)) : acos(-sqrt(up));   // kerr_raytracing_utils.py:257-262
        if (thminus) thminus[i] = acos(sqrt(up));
        if (uplus) uplus[i] = up;
        if (uminus) uminus[i] = um;
        if (uratio) uratio[i] = ratio;
        if (a2uminus) a2uminus[i] = a2um;
    }
    if (tau_tot) tau_tot[i] = tau_total_of(o, quartic_roots(o, lam, eta, (mode & M_POLISH_ROOTS) != 0));
}

__global__ void __launch_bounds__(SETUP_THREADS) mino_total_kernel(Observer o, size_t n, const double* __restrict__ roots,
                                                                   double* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Roots R;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        R.re[k] = roots[(size_t)k * n + i];
        R.im[k] = roots[(size_t)(4 + k) * n + i];
    }
    R.c12 = R.im[0] != 0.0;                 // mask_IV  (kerr_raytracing_utils.py:412)
    R.c34 = R.im[2] != 0.0;                 // mask_III (:445)
    R.region = 0;
    out[i] = tau_total_of(o, R);
}

// ---------------------------------------------------------------------------------------------------------------------
// the track kernel
// ---------------------------------------------------------------------------------------------------------------------
// Mapping of the track kernel: a CTA owns FOUR CONSECUTIVE RAYS and all its warps walk their Mino-time samples together:
// lane = 4 * (sample within the warp's 8-sample block) + (ray within the group); warp w takes blocks w, w + TRK_WARPS, ...
//  * the four lanes of one sample write four consecutive pixels of one row of the reference's (ngeo, npix) arrays = one full
//    32-byte sector, so every store instruction writes 8 whole sectors and DRAM never sees a partial one;
//  * branches on the root class, the polar class, the theta-vs-Carlson choice and the number of harmonics are properties of
//    the ray: a warp diverges only where its 4 neighbouring pixels straddle a class boundary (the first mapping, 32
//    neighbouring pixels per warp, paid the Carlson path for 32 rays whenever one of them needed it);
//  * all warps of the CTA evaluate the SAME rays at different Mino times, i.e. they run the same code at the same time:
//    the ~8 k-instruction kernel stays hot in the instruction caches (one-ray-group-per-warp variants lost 20 % to
//    instruction fetch stalls, profiles/);
//  * the ~90 per-ray constants a point reads come from a [slot][4] shared-memory block (4 distinct addresses per load:
//    one wavefront), staged once per CTA.
#ifndef KG_TRK_WARPS
#define KG_TRK_WARPS 8
#endif
#ifndef KG_TRK_MINBLOCKS
#define KG_TRK_MINBLOCKS 2
#endif
constexpr int TRK_WARPS = KG_TRK_WARPS;             // warps per CTA
constexpr int TRK_THREADS = 32 * TRK_WARPS;
constexpr int TRK_RPW = 4;                          // rays per warp = doubles per 32-byte sector
#ifndef KG_TRK_PAD_SMEM
#define KG_TRK_PAD_SMEM 0               // experiments only: unused shared memory that limits the CTAs per SM
#endif
constexpr size_t TRK_SMEM = (size_t)NW * TRK_RPW * sizeof(double) + KG_TRK_PAD_SMEM;

struct SmemW {
    const double* p;   // &sw[0][j]: slot i of ray j of this warp (a broadcast load)
    __device__ __forceinline__ double operator()(int i) const { return p[i * TRK_RPW]; }
};

struct TrackOut {
    double* t;
    double* r;
    double* th;
    double* ph;
    double* sig;
    double* tau;
};

#ifndef KG_TRK_GROUPS
#define KG_TRK_GROUPS 8
#endif
constexpr int TRK_GROUPS = KG_TRK_GROUPS;           // 4-ray groups a CTA works through one after the other

// Passes of a full (r, theta, phi, t) trace.  The Carlson-duplication fallback (0.1 % of the rays of BASELINE cfg 3) costs
// every thread 50-60 registers whether it runs or not (ptxas wants 154 for the complete per-point code, 64 without the
// fallback), so a full trace is two launches over the same tiles:
//   hot pass (track_hot_kernel)   rays that are `hot` (F_TH_THETA and F_R_THETA): theta-function code only, 64 registers,
//                                 32 warps / SM
//   P_FIXUP                       the other rays through the complete code, ONE RAY PER WARP STEP (lanes = 32 consecutive
//                                 samples): fix-up rays are isolated, with lane = 4 * sample + ray three lanes of four idle
//   P_ALL                         every ray through the complete code in one pass: do_phi_and_t = False, caller-supplied
//                                 Mino times (coords_at_tau), full = 2 (the A/B reference of the tests)
// Measured and dropped: splitting by four-ray GROUP instead (fix-up = every ray of a group with a fallback ray, so that both
// passes write whole 32-byte sectors): the fix-up pass then drags three theta-path rays through the Carlson code's divergent
// warps at 8 of 32 lanes -- 1.19 ms against 0.76 ms (profiles/r02_hot_variants.txt); polar and radial motion as separate
// passes of 61 / 79 registers -- 9.2 + 18.8 ms against 22.1 ms (profiles/r01_pass_variants.txt).
enum TrkPass { P_ALL = 0, P_FIXUP = 2 };

__device__ __forceinline__ bool trk_hot(int fl) { return (fl & (F_TH_THETA | F_R_THETA)) == (F_TH_THETA | F_R_THETA); }

// The flags of a tile's 32 rays, one per lane (every warp of the CTA loads them: one coalesced load), and the ballots of the
// rays inside the image and of the hot ones.  Uniform over the CTA.
struct TileRays {
    int flag;               // this lane's ray flag (lane = ray within the tile)
    unsigned in, hot;       // ray masks
};
__device__ __forceinline__ TileRays tile_rays(const int32_t* __restrict__ flags_in, size_t npix, int lane, int ngroups) {
    const size_t p = (size_t)blockIdx.x * (ngroups * TRK_RPW) + lane;
    const bool inside = lane < ngroups * TRK_RPW && p < npix;
    TileRays t;
    t.flag = inside ? __ldg(flags_in + p) : 0;
    t.in = __ballot_sync(0xffffffffu, inside);
    t.hot = __ballot_sync(0xffffffffu, inside && trk_hot(t.flag));
    return t;
}

template <bool FULL, int PASS, bool HAS_TAU>
__global__ void __launch_bounds__(TRK_THREADS, KG_TRK_MINBLOCKS)
track_kernel(Observer o, size_t npix, const double* __restrict__ w_in, const int32_t* __restrict__ flags_in, int ntau,
             double maxtaufrac, const double* __restrict__ tau_in, size_t tau_ld, TrackOut out, size_t ld, int k_begin,
             int k_end, int chunk) {
    constexpr int WARPS = TRK_WARPS, THREADS = 32 * WARPS;
    constexpr int KSTEP = WARPS * (32 / TRK_RPW);                        // samples one CTA-wide pass covers
    extern __shared__ double sw[];                                      // [NW][4] constants of the current ray group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // lane = 4 * (sample of the warp's 8-sample block) + ray.  (A 64-bit shared-memory load with 4 distinct addresses per half
    // warp costs two wavefronts, with <= 2 one -- research/ubench/lds_wavefronts.cu -- but the mapping that exploits this was
    // measured and did not pay: the LSU data pipe is not what the kernel waits for.)
    const int j = lane & (TRK_RPW - 1), ks = lane / TRK_RPW;
    // byte offset of this thread's ray column, passed through a shuffle: ptxas otherwise re-derives it from %tid (S2R, LOP3,
    // LEA: ~50 instructions per point) in front of the constant loads instead of keeping one register alive
    const unsigned joff = __shfl_sync(0xffffffffu, (unsigned)j * (unsigned)sizeof(double), lane);
    // this CTA covers samples [k0, k1) of the ntau-sample track; output row = k - k_begin
    const int k0 = k_begin + blockIdx.y * chunk;
    const int k1 = min(k_end, k0 + chunk);
    const int kfirst = k0 + warp * (32 / TRK_RPW) + ks;
    const int klast = (ntau > 1) ? ntau - 1 : -1;                        // numpy.linspace: the last sample is `stop` exactly
    static_assert(TRK_GROUPS * TRK_RPW <= 32, "one flag per lane");
    const TileRays tr = tile_rays(flags_in, npix, lane, TRK_GROUPS);
    const unsigned todo = (PASS == P_ALL) ? tr.in : (tr.in & ~tr.hot);  // rays of this pass
    if (todo == 0) return;                                              // uniform over the CTA: the fix-up pass is thousands of those
    bool staged = false;
    for (int g = 0; g < TRK_GROUPS; ++g) {
        const unsigned gm = (todo >> (g * TRK_RPW)) & ((1u << TRK_RPW) - 1u);
        if (gm == 0) continue;                                          // uniform over the CTA
        const size_t pix0 = ((size_t)blockIdx.x * TRK_GROUPS + g) * TRK_RPW;
        const int nvalid = (npix - pix0 < (size_t)TRK_RPW) ? (int)(npix - pix0) : TRK_RPW;
        const int fl = __shfl_sync(0xffffffffu, tr.flag, g * TRK_RPW + j);
        if (staged) __syncthreads();                                    // everyone is done with the previous group's constants
        staged = true;
        // the group's block of the workspace is [slot][4] already: a contiguous copy (rays beyond npix are padding)
        const double* const gsrc = w_in + (pix0 / TRK_RPW) * (size_t)kGroupDoubles;
        for (int i = threadIdx.x; i < NW * TRK_RPW; i += THREADS) sw[i] = gsrc[i];
        __syncthreads();
        if (PASS == P_FIXUP) {
            // a warp takes ONE ray at a time and its lanes 32 consecutive samples: the stores are 8-byte pieces of 32 different
            // rows -- partial sectors, but only for the 0.1 % of the rays traced here
            for (int jr = 0; jr < nvalid; ++jr) {
                if (!((gm >> jr) & 1u)) continue;                       // uniform over the CTA
                const int flr = __shfl_sync(0xffffffffu, tr.flag, g * TRK_RPW + jr);
                const size_t pixr = pix0 + jr;
                SmemW Wr{sw + jr};
                const double stop = maxtaufrac * Wr(S_TAUTOT);
                const double step = (ntau > 1) ? stop / (double)(ntau - 1) : 0.0;
                for (int k = k0 + warp * 32 + lane; k < k1; k += WARPS * 32) {
                    const double tau = (k == klast) ? stop : (double)k * step;
                    const Point p = point_eval<SmemW>(Wr, o, flr, tau, FULL);
                    const size_t q = (size_t)(k - k_begin) * ld + pixr;
                    if (out.r) __stcs(out.r + q, p.r);
                    if (out.th) __stcs(out.th + q, p.th);
                    if (out.t) __stcs(out.t + q, FULL ? p.t : 0.0);
                    if (out.ph) __stcs(out.ph + q, FULL ? p.ph : 0.0);
                    if (out.sig) __stcs(out.sig + q, FULL ? p.sig : 0.0);
                    if (out.tau) __stcs(out.tau + q, tau);
                }
            }
            continue;
        }
        if (j >= nvalid) continue;
        const size_t pix = pix0 + j;
        SmemW W{reinterpret_cast<const double*>(reinterpret_cast<const char*>(sw) + joff)};
        // numpy.linspace(0, stop, ntau): k * (stop / (ntau-1)), last sample = stop exactly
        const double stop = maxtaufrac * W(S_TAUTOT);
        const double step = (ntau > 1) ? stop / (double)(ntau - 1) : 0.0;
        size_t q = (size_t)(kfirst - k_begin) * ld + pix;              // output index, advanced by KSTEP rows per sample
        const size_t qstep = (size_t)KSTEP * ld;
        for (int k = kfirst; k < k1; k += KSTEP, q += qstep) {
            double tau;
            if (HAS_TAU) tau = tau_in[(size_t)(k - k_begin) * tau_ld + pix];
            else tau = (k == klast) ? stop : (double)k * step;
            const Point p = point_eval<SmemW>(W, o, fl, tau, FULL);
            if (out.r) __stcs(out.r + q, p.r);
            if (out.th) __stcs(out.th + q, p.th);
            if (out.t) __stcs(out.t + q, FULL ? p.t : 0.0);
            if (out.ph) __stcs(out.ph + q, FULL ? p.ph : 0.0);
            if (out.sig) __stcs(out.sig + q, FULL ? p.sig : 0.0);
            if (out.tau) __stcs(out.tau + q, tau);
        }
    }
}

// The hot pass of a full trace (rays with F_TH_THETA and F_R_THETA): the lane mapping of track_kernel (lane = 4 * sample + ray,
// whole 32-byte sectors per store) with the theta-function-only code of kg_hot.cuh.  The constants of a four-ray group are one
// contiguous block of the grouped workspace; they are staged into shared memory by a cooperative contiguous copy (two
// __syncthreads per group).  Every alternative to that staging was built and measured on cfg 3 (profiles/r02_hot_variants.txt;
// kept behind KG_HOT_TMA for A/B):
//   * 1-D TMA bulk copies (cp.async.bulk + mbarrier) of all 8 groups of the CTA up front: 20.1 against 19.3 ms -- 4 x 48 KB of
//     shared memory leave the L1 too small for the kernel's ~100 B/thread of register spills (local-load hit rate 23 %);
//   * the same through a ring of 2 / 3 buffers with full / empty mbarriers and an elected producer thread: 20.4 / 19.8 ms;
//   * no staging at all, the constants read with ld.global.nc from the L1-resident block: 21.9 ms (a 64-bit base instead of a
//     32-bit shared address: more spills, and the L1 tag stage per load);
//   * two / three Mino-time samples per thread (KG_HOT_NS: constants and ray-property branches paid once per NS samples, NS
//     independent dependency chains): 21.7 / 31.0 ms -- 128 / 168 registers halve / third the resident warps;
//   * 56 / 48 registers (36 / 40 warps per SM): 20.7 / 21.9 ms; 72 registers with 7-warp CTAs (28 warps): 19.8 ms.
// ncu of the shipped configuration: FP64 pipe 62 %, issue slots 66 %, LSU data pipe 76 % busy -- per 32 points ~200 cycles of
// FP64 pipe, ~215 of issue and ~230 of shared-memory wavefronts against 320 elapsed: three co-limiters within 15 % of each
// other, which is why removing instructions of one kind (8 DFMA per point, 8 local-memory round trips per point) moved nothing.
#ifndef KG_HOT_NS
#define KG_HOT_NS 1                     // Mino-time samples per thread (2 was measured: 21.7 against 19.3 ms, it halves the warps)
#endif
#ifndef KG_HOT_WARPS
#define KG_HOT_WARPS 8
#endif
#ifndef KG_HOT_MINBLOCKS
#define KG_HOT_MINBLOCKS 4
#endif
#ifndef KG_HOT_TMA
#define KG_HOT_TMA 0                    // 0: cooperative staging with two barriers per group (the round-1 scheme), for A/B
#endif
#ifndef KG_HOT_NBUF
#define KG_HOT_NBUF 2                   // ring of group buffers the bulk copies land in
#endif
constexpr int HOT_NS = KG_HOT_NS, HOT_WARPS = KG_HOT_WARPS;
constexpr int HOT_GROUPS = TRK_GROUPS, HOT_NBUF = KG_HOT_NBUF;
constexpr unsigned HOT_GROUP_BYTES = NW_HOT * TRK_RPW * sizeof(double);
static_assert(HOT_GROUP_BYTES % 16 == 0 && (kGroupDoubles * sizeof(double)) % 16 == 0, "bulk copies move multiples of 16 bytes");
constexpr size_t HOT_SMEM = KG_HOT_TMA ? (size_t)HOT_NBUF * HOT_GROUP_BYTES + 2 * HOT_NBUF * sizeof(unsigned long long)
                                       : (size_t)HOT_GROUP_BYTES;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "KG_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra KG_MBAR_DONE;\n"
        "bra KG_MBAR_WAIT;\n"
        "KG_MBAR_DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int NS>
__global__ void __launch_bounds__(32 * HOT_WARPS, KG_HOT_MINBLOCKS)
track_hot_kernel(Observer o, size_t npix, const double* __restrict__ w_in, const int32_t* __restrict__ flags_in, int ntau,
                 double maxtaufrac, TrackOut out, size_t ld, int k_begin, int k_end, int chunk) {
    constexpr int WARPS = HOT_WARPS, THREADS = 32 * WARPS;
    constexpr int KSTEP = WARPS * (32 / TRK_RPW);                        // samples one CTA-wide pass covers (per sample slot)
    extern __shared__ __align__(128) unsigned char hot_smem[];
    double* const sw = reinterpret_cast<double*>(hot_smem);             // [HOT_GROUPS][NW_HOT][4] (TMA) or [NW_HOT][4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane & (TRK_RPW - 1), ks = lane / TRK_RPW;
    // byte offset of this thread's ray column, passed through a shuffle so that ptxas keeps it in a register (see track_kernel)
    const unsigned joff = __shfl_sync(0xffffffffu, (unsigned)j * (unsigned)sizeof(double), lane);
    const int k0 = k_begin + blockIdx.y * chunk;
    const int k1 = min(k_end, k0 + chunk);
    const int kfirst = k0 + warp * (32 / TRK_RPW) + ks;
    const int klast = (ntau > 1) ? ntau - 1 : -1;                        // numpy.linspace: the last sample is `stop` exactly
    static_assert(HOT_GROUPS * TRK_RPW <= 32, "one flag per lane");
    const TileRays tr = tile_rays(flags_in, npix, lane, HOT_GROUPS);
    if (tr.hot == 0) return;                                            // uniform over the CTA
    unsigned gmask = 0;                                                 // groups of this tile with hot rays
#pragma unroll
    for (int g = 0; g < HOT_GROUPS; ++g)
        if ((tr.hot >> (g * TRK_RPW)) & ((1u << TRK_RPW) - 1u)) gmask |= 1u << g;
    const double* const tile_src = w_in + (size_t)blockIdx.x * HOT_GROUPS * (size_t)kGroupDoubles;
#if KG_HOT_TMA
    unsigned long long* const full = reinterpret_cast<unsigned long long*>(hot_smem + (size_t)HOT_NBUF * HOT_GROUP_BYTES);
    unsigned long long* const empty = full + HOT_NBUF;
    // the n-th hot group of the tile (bit order of gmask) uses ring buffer n % HOT_NBUF
    auto issue = [&](int g, int buf) {
        mbar_expect_tx(full + buf, HOT_GROUP_BYTES);
        bulk_copy_g2s(hot_smem + (size_t)buf * HOT_GROUP_BYTES, tile_src + (size_t)g * kGroupDoubles, HOT_GROUP_BYTES, full + buf);
    };
    unsigned pending = gmask;                 // groups not yet requested (only thread 0 uses it)
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < HOT_NBUF; ++b) {
            mbar_init(full + b, 1);
            mbar_init(empty + b, WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // inits visible to the async proxy
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int b = 0; b < HOT_NBUF && pending; ++b) {
            const int g = __ffs(pending) - 1;
            pending &= pending - 1;
            issue(g, b);
        }
    }
    int nth = 0;                              // index of the current group among the tile's hot groups
#else
    bool staged = false;
#endif
    for (int g = 0; g < HOT_GROUPS; ++g) {
        if (!((gmask >> g) & 1u)) continue;                             // uniform over the CTA
        const size_t pix0 = ((size_t)blockIdx.x * HOT_GROUPS + g) * TRK_RPW;
        const int nvalid = (npix - pix0 < (size_t)TRK_RPW) ? (int)(npix - pix0) : TRK_RPW;
        const int fl = __shfl_sync(0xffffffffu, tr.flag, g * TRK_RPW + j);
#if KG_HOT_TMA
        const int buf = nth % HOT_NBUF;
        const unsigned round = (unsigned)(nth / HOT_NBUF);
        constexpr int LAG = HOT_NBUF - 1;
        if (nth >= LAG && threadIdx.x == 0 && pending) {
            // refill the buffer of the group traced LAG groups ago (every warp arrives on its `empty` barrier when it is done
            // with a group) with the next group not yet requested, i.e. the one after the current: it lands while the current
            // group is traced.  With two buffers this warp waits here for the slowest warp of the previous group.
            const int pbuf = (nth - LAG) % HOT_NBUF;
            mbar_wait(empty + pbuf, (unsigned)((nth - LAG) / HOT_NBUF) & 1u);
            const int gn = __ffs(pending) - 1;
            pending &= pending - 1;
            issue(gn, pbuf);
        }
        ++nth;
        mbar_wait(full + buf, round & 1u);
        const double* const gsw = sw + (size_t)buf * (NW_HOT * TRK_RPW);
#else
        if (staged) __syncthreads();                                    // everyone is done with the previous group's constants
        staged = true;
        for (int i = threadIdx.x; i < NW_HOT * TRK_RPW; i += THREADS) sw[i] = tile_src[(size_t)g * kGroupDoubles + i];
        __syncthreads();
        const double* const gsw = sw;
#endif
        if (j < nvalid && trk_hot(fl)) {
        const size_t pix = pix0 + j;
        using HotW = SmemW;
        HotW W{reinterpret_cast<const double*>(reinterpret_cast<const char*>(gsw) + joff)};
        // numpy.linspace(0, stop, ntau): k * (stop / (ntau-1)), last sample = stop exactly
        const double stop = maxtaufrac * W(S_TAUTOT);
        const double step = (ntau > 1) ? stop / (double)(ntau - 1) : 0.0;
        const double lam = W(S_LAM);
        size_t q = (size_t)(kfirst - k_begin) * ld + pix;              // output index of sample slot 0
        const size_t qs = (size_t)KSTEP * ld, qstep = (size_t)NS * qs;
        for (int k = kfirst; k < k1; k += NS * KSTEP, q += qstep) {
            double tau[NS];
            bool live[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int kk = k + s * KSTEP;
                live[s] = (s == 0) || kk < k1;                          // a dead slot re-evaluates sample k and stores nothing
                const int kc = live[s] ? kk : k;
                tau[s] = (kc == klast) ? stop : (double)kc * step;
            }
            double th[NS], Gph[NS], Gt[NS], r[NS], Iph[NS], It[NS], Isig[NS];
            polar_hot<HotW, NS>(W, fl, tau, th, Gph, Gt);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (out.th && live[s]) __stcs(out.th + q + s * qs, th[s]);
                Gph[s] *= lam;
                Gt[s] *= o.a2;
            }
            radial_hot<HotW, NS>(W, o, fl, tau, r, Iph, It, Isig);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (!live[s]) continue;
                const size_t qq = q + s * qs;
                if (out.r) __stcs(out.r + qq, r[s]);
                if (out.t) __stcs(out.t + qq, It[s] + Gt[s]);
                if (out.ph) __stcs(out.ph + qq, Iph[s] + Gph[s]);
                if (out.sig) __stcs(out.sig + qq, Isig[s] + Gt[s]);
                if (out.tau) __stcs(out.tau + qq, tau[s]);
            }
        }
        }
#if KG_HOT_TMA
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + buf);                        // this warp is done with the group's constants
#endif
    }
}

__global__ void __launch_bounds__(TRK_THREADS) crossing_counts_kernel(Observer o, size_t npix, const double* __restrict__ beta,
                                                                      const double* __restrict__ w_in,
                                                                      const int32_t* __restrict__ flags_in, int ntau,
                                                                      const double* __restrict__ tau_in, double* neq,
                                                                      double* npol) {
    size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const StridedW W = ray_view(w_in, pix);
    int fl = flags_in[pix];
    double b = beta[pix];
    for (int k = blockIdx.y; k < ntau; k += gridDim.y) {
        size_t q = (size_t)k * npix + pix;
        double tau = tau_in[q];
        if (neq) neq[q] = n_equatorial_at(W, o, fl, b, tau);
        if (npol) npol[q] = n_poloidal_at(W, tau);
    }
}

// th_integrate / r_integrate as separate calls (kgeo/kerr_raytracing_ana.py:195-365, 368-633): rays set up from the conserved
// quantities, then the angular antiderivatives (theta, G_phi, G_t) and / or the radial ones (r, I_phi, I_t, I_sigma) at the user's
// Mino times through the complete per-point code.  One thread per ray, grid.y over the samples.
template <bool PHI_T>
__global__ void __launch_bounds__(SETUP_THREADS, KG_SETUP_MINBLOCKS) ray_setup_conserved_kernel(Observer o, size_t npix, const double* __restrict__ lam,
                                                                            const double* __restrict__ eta, const double* __restrict__ s_o,
                                                                            int mode, double* __restrict__ w_out,
                                                                            int32_t* __restrict__ flags_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    int fl;
    ray_setup_conserved<PHI_T, true, true>(o, lam[i], eta[i], s_o ? (s_o[i] >= 0.0) : true, mode, ray_ptr(w_out, i), &fl);
    flags_out[i] = fl;
}

struct IntegralsOut {
    double *th, *Gph, *Gt, *r, *Iph, *It, *Isig;
};
__global__ void __launch_bounds__(TRK_THREADS) track_integrals_kernel(Observer o, size_t npix, const double* __restrict__ w_in,
                                                                      const int32_t* __restrict__ flags_in, int ntau,
                                                                      const double* __restrict__ tau_in, size_t tau_ld, bool full,
                                                                      IntegralsOut out, size_t ld) {
    size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const StridedW W = ray_view(w_in, pix);
    const int fl = flags_in[pix];
    const bool polar = out.th || out.Gph || out.Gt, radial = out.r || out.Iph || out.It || out.Isig;
    for (int k = blockIdx.y; k < ntau; k += gridDim.y) {
        const double tau = tau_in[(size_t)k * tau_ld + pix];
        const size_t q = (size_t)k * ld + pix;
        if (polar) {
            double th, Gph, Gt;
            theta_eval<StridedW>(W, fl, tau, full, &th, &Gph, &Gt);
            if (out.th) out.th[q] = th;
            if (out.Gph) out.Gph[q] = full ? Gph : 0.0;
            if (out.Gt) out.Gt[q] = full ? Gt : 0.0;
        }
        if (radial) {
            double r, Iph, It, Isig;
            radial_eval<StridedW, true>(W, o, fl, tau, full, &r, &Iph, &It, &Isig);
            if (out.r) out.r[q] = r;
            if (out.Iph) out.Iph[q] = full ? Iph : 0.0;
            if (out.It) out.It[q] = full ? It : -4.0 * tau;      // the reference's I_t = I_tA + 4 I_0 + 2 I_1 + I_2 keeps I_0 = -tau (:627-631)
            if (out.Isig) out.Isig[q] = full ? Isig : 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// field-line crossings (kgeo/solver.py: getguesses / findroot)
// ---------------------------------------------------------------------------------------------------------------------
struct PsiSpec {
    int model;
    double pval, psitarget;
};

// getguesses (kgeo/solver.py:98-137), one thread per ray.  PASS 0 counts the ray's guesses (and takes the maximum over the
// rays), PASS 1 writes them into guesses[m * npix + ray] (padded with -1: makegoodarray + transpose).
template <int PASS>
__global__ void __launch_bounds__(256) fieldline_guess_kernel(Observer o, PsiSpec ps, size_t npix, int ngeo,
                                                              const double* __restrict__ r, const double* __restrict__ th,
                                                              const double* __restrict__ tau, size_t ld,
                                                              const double* __restrict__ alpha, const double* __restrict__ beta,
                                                              const double* __restrict__ w_in, const int32_t* __restrict__ flags_in,
                                                              int neqmax, double maxtaufrac, int32_t* count, int32_t* maxcount,
                                                              int maxlen, double* guesses, double* taumax_out) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const StridedW W = ray_view(w_in, pix);
    const int fl = flags_in[pix];
    const double al = alpha[pix], be = beta[pix];
    const double taumax = maxtau_forwardjet(W, fl, be, neqmax, maxtaufrac);
    if (PASS == 0 && taumax_out) taumax_out[pix] = taumax;
    int n = 0;
    auto emit = [&](double g) {
        if (PASS == 1 && n < maxlen) guesses[(size_t)n * npix + pix] = g;
        ++n;
    };
    const double impact = sqrt(al * al + be * be);
    bool flat = false;
    if (fabs(o.sth) < 0.1 && impact > 100.0) {                          // flat-space face-on guesses
        double rparam = 0.0;
        if (ps.model == PSI_POWER && ps.pval > 0.0) {
            rparam = pow(impact * impact / (2.0 * ps.psitarget), 1.0 / (2.0 - ps.pval));
            flat = true;
        } else if (ps.model == PSI_PARA) {
            rparam = (impact * impact + ps.psitarget * ps.psitarget) / (2.0 * ps.psitarget);
            flat = true;
        }
        if (flat) {
            emit(1.0 / rparam);
            emit(2.0 / impact - 1.0 / rparam);
        }
    }
    if (!flat && ngeo > 0) {
        // indmax = argmin_k |tau_k - taumax| (first minimum; NaN distances never win, all-NaN gives 0 as np.argmin does)
        int indmax = 0;
        double best = fabs(tau[pix] - taumax);
        for (int k = 1; k < ngeo; ++k) {
            const double d = fabs(tau[(size_t)k * ld + pix] - taumax);
            if (best != best) {
                // np.argmin returns the first NaN: keep index of the first NaN
            } else if (d != d || d < best) {
                best = d;
                indmax = k;
                if (d != d) break;
            }
        }
        // minfunc[k] = psi[k+1] psi[k] on k = 1 .. indmax-2 ("don't trust left or right endpoint")
        double pk = 0.0;
        bool have = false;
        for (int k = 1; k <= indmax - 2; ++k) {
            const size_t q = (size_t)k * ld + pix;
            if (!have) pk = psi_value(ps.model, ps.pval, o.a, r[q], th[q]) - ps.psitarget;
            const size_t q1 = (size_t)((k + 1) % ngeo) * ld + pix;
            const double pn = psi_value(ps.model, ps.pval, o.a, r[q1], th[q1]) - ps.psitarget;
            if (pn * pk < 0.0) emit(0.5 * (tau[q] + tau[q1]));
            pk = pn;
            have = true;
        }
    }
    if (n == 0) emit(-1.0);
    if (PASS == 0) {
        count[pix] = n;
        atomicMax(maxcount, n);
    } else {
        for (int m = n; m < maxlen; ++m) guesses[(size_t)m * npix + pix] = -1.0;
    }
}

// state of scipy.optimize.newton's vectorised secant iteration (_array_newton), element e = m * npix + ray
struct SecantState {
    double *p, *p1, *q0, *q1, *pnew;
    int32_t *active, *fail;
    int32_t* gflags;    // [0] any(q1 != q0), [1] any(failures[nz_der])
};

__global__ void __launch_bounds__(256) fieldline_eval_kernel(Observer o, PsiSpec ps, size_t npix, size_t n,
                                                             const double* __restrict__ w_in, const int32_t* __restrict__ flags_in,
                                                             const double* __restrict__ guesses, const double* __restrict__ tau,
                                                             double* q, double* r_out, double* th_out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const size_t pix = e % npix;
    const StridedW W = ray_view(w_in, pix);
    double rr, tt;
    double v = psi_residual(W, o, flags_in[pix], ps.model, ps.pval, ps.psitarget, tau[e], &rr, &tt);
    if (guesses[e] == -1.0) v = 0.0;                                    // arrhere[guesses == -1] = 0
    if (q) q[e] = v;
    if (r_out) r_out[e] = rr;
    if (th_out) th_out[e] = tt;
}

__global__ void __launch_bounds__(256) fieldline_secant_start_kernel(size_t n, const double* __restrict__ guesses, SecantState st) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double dx = 6.828499381469512e-06;                            // np.finfo(float).eps ** 0.33
    const double p = guesses[e];
    st.p[e] = p;
    st.p1[e] = p * (1.0 + dx) + ((p >= 0.0) ? dx : -dx);
    st.active[e] = 1;
    st.fail[e] = 1;
}

__global__ void __launch_bounds__(256) fieldline_secant_step_kernel(size_t n, double tol, SecantState st) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double p = st.p[e], p1 = st.p1[e], q0 = st.q0[e], q1 = st.q1[e];
    const bool nz = (q1 != q0);
    double pn = p;
    int act = st.active[e], fail = st.fail[e] & 1;
    if (nz) {
        const double dp = (q1 * (p1 - p)) / (q1 - q0);
        pn = p1 - dp;
        fail = (fabs(dp) >= tol) ? 1 : 0;
        if (fail) atomicOr(st.gflags + 1, 1);
        atomicOr(st.gflags + 0, 1);
    } else if (act) {
        pn = (p1 + p) / 2.0;
    }
    act = act && nz;
    st.pnew[e] = pn;
    st.active[e] = act;
    st.fail[e] = (fail & 1) | (nz ? 0 : 2);           // bit 0: scipy's `failures`, bit 1: zero derivative at this step
}

__global__ void __launch_bounds__(256) fieldline_average_kernel(size_t n, const double* p, const double* p1, double* out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = (p1[e] + p[e]) / 2.0;
}

// tail of findroot (kgeo/solver.py:181-207): coordinates at the root and at (1 + 1e-5) times it, momentum signs, crossing count
__global__ void __launch_bounds__(256) fieldline_finish_kernel(Observer o, size_t npix, size_t n, const double* __restrict__ w_in,
                                                               const int32_t* __restrict__ flags_in, const double* __restrict__ beta,
                                                               const double* __restrict__ guesses, const double* __restrict__ root,
                                                               double* tau_out, double* r_out, double* th_out, double* signpr,
                                                               double* signpth, double* neq) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const size_t pix = e % npix;
    const StridedW W = ray_view(w_in, pix);
    const int fl = flags_in[pix];
    const double tau = root[e];
    const Point a = point_eval(W, o, fl, tau, false);
    const Point b = point_eval(W, o, fl, tau * (1.0 + 1.0e-5), false);
    const bool none = (guesses[e] == -1.0);
    if (neq) neq[e] = neq_of_tau(W, fl, beta[pix], tau);
    tau_out[e] = none ? 0.0 : tau;
    r_out[e] = none ? 0.0 : a.r;
    th_out[e] = none ? 0.0 : a.th;
    signpr[e] = none ? 0.0 : np_sign(a.r - b.r);
    signpth[e] = none ? 0.0 : np_sign(a.th - b.th);
}

// ---------------------------------------------------------------------------------------------------------------------
// off-equatorial jet emission (kgeo/off_eq.py getstokes: density + Iobs_off per crossing, geometry.sort_image)
// The per-crossing device functions are also checked on the CPU through a host build (tests/test_offeq_hostbuild.py); the
// kernels against the unmodified reference's getstokes in tests/test_gpu_jet.py.
// ---------------------------------------------------------------------------------------------------------------------
KG_HD double nan_to_num(double x) {
    if (x != x) return 0.0;
    if (x > 1.7976931348623157e308) return 1.7976931348623157e308;
    if (x < -1.7976931348623157e308) return -1.7976931348623157e308;
    return x;
}

// Boost of one crossing: the velocity of Iobs_off takes ONE field line's (r0, omega, E - Omega_F L) for all crossings (`glob`,
// from the first crossing of the flattened arrays, kgeo/velocities.py:495-516), the velocity inside the density takes one per
// crossing ORDER (row m of the (nguess, npix) arrays: density_*_all loop over the rows, kgeo/densityfuncs.py:60-72).
struct JetBoostRows {
    const double* rows;     // [nguess][3] = r0, omega, Aconst per crossing order (FF only), or NULL
};
__device__ __forceinline__ JetBoost density_boost(const JetBoost& glob, const JetBoostRows& R, size_t m) {
    JetBoost b = glob;
    b.gammamax = 0.0;       // gammamax = None in density_*_all
    b.gscale = 1.0;
    if (glob.ff && R.rows) {
        b.r0 = R.rows[3 * m];
        b.omega = R.rows[3 * m + 1];
        b.Aconst = R.rows[3 * m + 2];
    }
    return b;
}

// |nan_to_num(x, nan = inf)| as NumPy computes it (+-inf become +-DBL_MAX), as the bits of a non-negative double (ordered like
// unsigned integers)
__device__ __forceinline__ unsigned long long argmin_key(double x) {
    double v;
    if (x != x) v = KG_INF;
    else if (isinf(x)) v = 1.7976931348623157e308;
    else v = fabs(x);
    return (unsigned long long)__double_as_longlong(v);
}

// gammamax pre-pass (kgeo/velocities.py:593-603): gamma_FF and gamma_eff of every crossing, and the smallest |gamma_eff|
__global__ void __launch_bounds__(256) jet_gamma_kernel(JetSpec J, double a, size_t n, const double* __restrict__ r,
                                                        const double* __restrict__ th, JetBoost boost, double* gam, double* geff,
                                                        unsigned long long* minkey) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    boost.gscale = 1.0;
    const DriftOut d = jet_driftframe_boost(J, a, r[e], th[e], boost);
    gam[e] = d.gamma;
    geff[e] = d.gammaeff;
    atomicMin(minkey, argmin_key(d.gammaeff));
}
// np.argmin: the FIRST element that attains the minimum
__global__ void __launch_bounds__(256) jet_argmin_kernel(size_t n, const double* __restrict__ geff, const unsigned long long* minkey,
                                                         unsigned long long* argidx) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (argmin_key(geff[e]) == *minkey) atomicMin(argidx, (unsigned long long)e);
}
__global__ void jet_gscale_kernel(const double* gam, const double* geff, const unsigned long long* argidx, double* gscale) {
    const unsigned long long i = *argidx;
    gscale[0] = gam[i] / geff[i];
}

__global__ void __launch_bounds__(256) jet_stokes_kernel(JetSpec J, double a, double th_o, size_t npix, size_t n,
                                                         const double* __restrict__ alpha, const double* __restrict__ beta,
                                                         const double* __restrict__ r, const double* __restrict__ th,
                                                         const double* __restrict__ kr, const double* __restrict__ kth,
                                                         JetBoost boost, JetBoostRows rows, const double* __restrict__ gscale,
                                                         double sigma, double r0, double norm, double psitarget, double specind,
                                                         double* I, double* Q, double* U, double* sinthb, double* dens,
                                                         int32_t* nu_out_of_range) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const size_t pix = e % npix;
    const double rs = r[e], ths = th[e];
    if (gscale) boost.gscale = gscale[0];
    double nu_d = 0.0;
    const double rho = jet_density(J, a, rs, ths, density_boost(boost, rows, e / npix), sigma, r0, norm, psitarget, &nu_d);
    if (dens) dens[e] = rho;
    const StokesOut s = iobs_off_point(J, a, th_o, alpha[pix], beta[pix], rs, ths, kr[e], kth[e], boost, rho, specind);
    // u_driftframe raises when any boost leaves [-1, 1] (NaN compares false, as in NumPy): reported to the host wrapper
    if (nu_out_of_range && (s.nu > 1.0 || s.nu < -1.0 || nu_d > 1.0 || nu_d < -1.0)) atomicOr(nu_out_of_range, 1);
    const bool none = (rs == 0.0);                                       // iobs[rvals == 0] = 0: no field-line crossing
    I[e] = none ? 0.0 : nan_to_num(s.I);
    Q[e] = none ? 0.0 : nan_to_num(s.Q);
    U[e] = none ? 0.0 : nan_to_num(s.U);
    if (sinthb) sinthb[e] = nan_to_num(s.sinthb);
}

// geometry.sort_image (kgeo/geometry.py:38-62): planes 0..neqmax-1 = sum of the crossings of that order, plane neqmax = all
__global__ void __launch_bounds__(256) jet_sort_kernel(size_t npix, int nguess, int neqmax, const double* __restrict__ I,
                                                       const double* __restrict__ Q, const double* __restrict__ U,
                                                       const double* __restrict__ neq, double* oI, double* oQ, double* oU) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    for (int plane = 0; plane <= neqmax; ++plane) {
        double si = 0.0, sq_ = 0.0, su = 0.0;
        for (int m = 0; m < nguess; ++m) {
            const size_t e = (size_t)m * npix + pix;
            if (plane == neqmax || neq[e] == (double)plane) {
                si += I[e];
                sq_ += Q[e];
                su += U[e];
            } else {
                si += 0.0; sq_ += 0.0; su += 0.0;
            }
        }
        oI[(size_t)plane * npix + pix] = si;
        oQ[(size_t)plane * npix + pix] = sq_;
        oU[(size_t)plane * npix + pix] = su;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// equatorial lensing / image
// ---------------------------------------------------------------------------------------------------------------------
#ifndef KG_IMG_THREADS
#define KG_IMG_THREADS 384
#endif
#ifndef KG_IMG_MINBLOCKS
#define KG_IMG_MINBLOCKS 2
#endif
constexpr int IMG_THREADS = KG_IMG_THREADS;

__global__ void __launch_bounds__(IMG_THREADS, KG_IMG_MINBLOCKS) r_equatorial_kernel(Observer o, size_t npix, const double* __restrict__ alpha,
                                                                   const double* __restrict__ beta, int mbar_lo, int mbar_hi,
                                                                   double* rs_out, double* Ir_out, double* Nmax_out,
                                                                   double* Imax_out, double* npol_out, unsigned char* outside_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    double w[NW_LITE];
    int fl;
    double be = beta[i];
    ray_setup<false, false>(o, alpha[i], be, M_ANY_ETA_POS, w, &fl);
    ArrayW W{w};
    const bool reg = w[S_ETA] > 0.0;
    if (outside_out) outside_out[i] = (fl & F_OUTSIDE) ? 1 : 0;      // is_outside_crit: the emission kernel's radial momentum sign
    if (Imax_out) Imax_out[i] = reg ? w[S_TAUTOT] : -1.0;
    if (npol_out) npol_out[i] = n_poloidal_at(W, w[S_TAUTOT]);
    for (int mbar = mbar_lo; mbar <= mbar_hi; ++mbar) {
        double rs = -1.0, Ir = -1.0, Nmax = -2.0;
        if (reg) {
            crossing_time(W, o, be, mbar, &Ir, &Nmax);
            double d0, d1, d2;
            radial_eval<ArrayW, false>(W, o, fl, Ir, false, &rs, &d0, &d1, &d2);
            if (Nmax < (double)mbar) rs = 0.0;
        }
        size_t q = (size_t)(mbar - mbar_lo) * npix + i;
        if (rs_out) rs_out[q] = rs;
        if (Ir_out) Ir_out[q] = Ir;
        if (Nmax_out) Nmax_out[q] = Nmax;
    }
}

// ---- inverse lensing: rho_of_req (kgeo/equatorial_lensing.py:197-266) ----------------------------------------------------
// objective of the root find: r_equatorial2(rho, varphi) - req (:197-224), observer at infinity
__device__ __noinline__ double req_objective(const Observer& o, double rho, double cv, double sv, int mbar, double req) {
    double w[NW_LITE];
    int fl;
    const double al = cv * rho, be = sv * rho;
    ray_setup<false, false>(o, al, be, M_ANY_ETA_POS, w, &fl);
    ArrayW W{w};
    double rs = -1.0;
    if (w[S_ETA] > 0.0) {
        double Ir, Nmax, d0, d1, d2;
        crossing_time(W, o, be, mbar, &Ir, &Nmax);
        radial_eval<ArrayW, false>(W, o, fl, Ir, false, &rs, &d0, &d1, &d2);
        if (Nmax < (double)mbar) rs = 0.0;
    }
    if (rs <= 0.0) rs = (fl & F_OUTSIDE) ? 1.0e50 : 1.0;     // :208-215 ("fudge to make things continuous")
    return rs - req;
}

// SciPy's brentq (scipy/optimize/Zeros/brentq.c), statement for statement, on rho -> r_equatorial2(rho) - req along the image-plane
// direction (cv, sv).  status: 0 converged, -1 no sign change over [xa, xb], -2 not converged after maxiter iterations.
__device__ double brent_rho(const Observer& o, double cv, double sv, int mbar, double req, double xa, double xb, double xtol,
                            double rtol, int maxiter, int* status_out) {
    double xpre = xa, xcur = xb, xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
    double fpre = req_objective(o, xpre, cv, sv, mbar, req);
    double fcur = req_objective(o, xcur, cv, sv, mbar, req);
    int status = -2;
    double root = 0.0;
    bool done = false;
    if (fpre == 0.0) { root = xpre; status = 0; done = true; }
    else if (fcur == 0.0) { root = xcur; status = 0; done = true; }
    else if (signbit(fpre) == signbit(fcur)) { root = 0.0; status = -1; done = true; }
    for (int it = 0; it < maxiter && !done; ++it) {
        if (fpre != 0.0 && fcur != 0.0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2;
        const double sbis = (xblk - xcur) / 2;
        if (fcur == 0.0 || fabs(sbis) < delta) {
            root = xcur; status = 0; done = true;
            break;
        }
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);                      // interpolate
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);                  // extrapolate
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
                spre = scur;                                                       // good short step
                scur = stry;
            } else {
                spre = sbis;                                                       // bisect
                scur = sbis;
            }
        } else {
            spre = sbis;                                                           // bisect
            scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = req_objective(o, xcur, cv, sv, mbar, req);
    }
    if (!done) root = xcur;
    *status_out = status;
    return root;
}

// one thread per angle
__global__ void __launch_bounds__(32) rho_of_req_kernel(Observer o, size_t n, const double* __restrict__ varphi, double req, int mbar,
                                                        double xa, double xb, double xtol, double rtol, int maxiter,
                                                        double* rho_out, double* alpha_out, double* beta_out,
                                                        int32_t* status_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cv = cos(varphi[i]), sv = sin(varphi[i]);
    int status;
    const double root = brent_rho(o, cv, sv, mbar, req, xa, xb, xtol, rtol, maxiter, &status);
    if (rho_out) rho_out[i] = root;
    if (alpha_out) alpha_out[i] = cv * root;
    if (beta_out) beta_out[i] = sv * root;
    if (status_out) status_out[i] = status;
}

// A whole library of curves in one launch (the reference's generate_library / make_curves2, kgeo/equatorial_lensing.py:351-417:
// a multiprocessing pool over (spin, inclination) configurations, each running rho_of_req for mbar = 0..mmax): one thread per
// (configuration, order, angle), the bracket of each order as rho_of_req picks it (:232-241).
__global__ void __launch_bounds__(32) rho_library_kernel(size_t nconf, const double* __restrict__ spin, const double* __restrict__ th0,
                                                         const double* __restrict__ req, int mbar_lo, int norder, size_t nphi,
                                                         const double* __restrict__ varphi, double xtol, double rtol, int maxiter,
                                                         double* __restrict__ rho_out, int32_t* __restrict__ status_out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nconf * norder * nphi) return;
    const size_t c = e / ((size_t)norder * nphi);
    const int mbar = mbar_lo + (int)((e / nphi) % norder);
    const size_t i = e % nphi;
    const Observer o = make_observer(spin[c], kInf, th0[c]);
    const double rq = req[c];
    const double xa = (mbar <= 1) ? 0.0 : 1.0;
    const double xb = (mbar == 0) ? rq + 5.0 : (mbar == 1 ? 100.0 : 10.0);
    int status;
    const double root = brent_rho(o, cos(varphi[i]), sin(varphi[i]), mbar, rq, xa, xb, xtol, rtol, maxiter, &status);
    rho_out[e] = root;
    if (status_out) status_out[e] = status;
}

// critical curve (kgeo/equatorial_lensing.py:268-290): NumPy's operation order with individually rounded operations (the
// square root of a difference that vanishes at both end points decides which entries are NaN)
__global__ void critical_curve_kernel(double a, double rminus, double rplus, double sth, double cth, double tth, size_t n,
                                      double* alpha, double* beta) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double step = (n > 1) ? div_rn(add_rn(rplus, -rminus), (double)(n - 1)) : 0.0;
    const double r = (i == n - 1 && n > 1) ? rplus : add_rn(rminus, mul_rn((double)i, step));     // numpy.linspace
    const double a2 = mul_rn(a, a), r2 = mul_rn(r, r);
    // r**3 as NumPy computes it (libm pow, correctly rounded in practice): r2 * r with the rounding error of r2 carried along
    const double r3 = fma(fma(r, r, -r2), r, mul_rn(r2, r)) ;
    const double num = add_rn(add_rn(add_rn(r3, a2), mul_rn(r, a2)), -mul_rn(3.0, r2));
    const double lam = div_rn(-num, add_rn(mul_rn(a, r), -a));
    const double eta = div_rn(mul_rn(r2, add_rn(add_rn(mul_rn(3.0, r2), a2), -mul_rn(lam, lam))), add_rn(r2, -a2));
    const double ac = mul_rn(a, cth), lt = div_rn(lam, tth);
    alpha[i] = div_rn(-lam, sth);
    beta[i] = sqrt_rn(add_rn(add_rn(eta, mul_rn(ac, ac)), -mul_rn(lt, lt)));
}

struct ImageOut {
    double *I, *Q, *U, *r, *Ir, *g, *sinthb, *nmax, *npol, *imax, *lp, *Ie;
};

// Equatorial image = TWO kernels.  r_equatorial_kernel (above) is the GEOMETRY: per pixel one setup, then per order the crossing
// time, Nmax and the crossing radius, written as planes.  image_emission_kernel is the EMISSION: per pixel and order the masks,
// momentum signs, fluid velocity, redshift, (polarization,) emissivity and Stokes parameters from the crossing planes.
// One fused kernel (rounds 1 and 2a: geometry of all orders parked in shared memory, then emission) executed only 35 % more
// instructions than the geometry alone but took 80 % longer (6.06 against 3.38 ms at 4096^2): ~17 k SASS instructions walked
// once per warp stalled on instruction fetch (no-instruction 3.8 per issue against 1.3) and on spill reloads around the
// 80-register cap.  Split, each kernel's code fits the instruction caches; the crossing planes r, Ir are outputs anyway and
// only Nmax (per order) and a byte per pixel travel through scratch.
struct GeomPlanes {
    const double *r, *Ir, *N;          // [norder, npix]
    const double *tau, *npol;          // [npix]: total Mino time (-1: vortical), nmax_poloidal (may be NULL)
    const unsigned char* outside;      // [npix]
};
#ifndef KG_EMIS_THREADS
#define KG_EMIS_THREADS 256
#endif
#ifndef KG_EMIS_MINBLOCKS
#define KG_EMIS_MINBLOCKS 4
#endif
constexpr int EMIS_THREADS = KG_EMIS_THREADS;
// SHARDED: the pixels are a block of image rows whose results go to rows lr * row_step + row_offset of destination planes that
// may live on ANOTHER GPU (peer memory over NVLink: plain coalesced stores -- the gather of a row-sharded image is fused into
// the kernel that computes it); the crossing planes are then copied along.
template <bool POL, bool SHARDED>
__global__ void __launch_bounds__(EMIS_THREADS, KG_EMIS_MINBLOCKS) image_emission_kernel(Observer o, kgeo_models M, kgeo_image_opts opt, size_t npix,
                                                            const double* __restrict__ alpha, const double* __restrict__ beta,
                                                            GeomPlanes g, ImageOut out, kgeo_shard_map map) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const size_t d = SHARDED ? ((i / map.cols) * map.row_step + map.row_offset) * map.cols + i % map.cols : i;
    const size_t pstride = SHARDED ? (size_t)map.plane_stride : npix;
    const double al = alpha[i], be = beta[i];
    const double lam = -al * o.sth, eta = conserved_eta(o, al, be);    // as ray_setup forms them
    const double tau_tot = g.tau[i];
    const int fl = g.outside[i] ? F_OUTSIDE : 0;
    if (out.npol && out.npol != g.npol) out.npol[d] = g.npol[i];
    if (out.imax && out.imax != g.tau) out.imax[d] = tau_tot;
    double lastN = -2.0;
#pragma unroll 1
    for (int mbar = opt.mbar_min; mbar <= opt.mbar_max; ++mbar) {
        const size_t qi = (size_t)(mbar - opt.mbar_min) * npix + i;
        const OrderGeom gq{g.r[qi], g.Ir[qi], g.N[qi]};
        const OrderOut v = image_physics<POL>(o, fl, M, opt, al, be, lam, eta, tau_tot, mbar, gq);
        const size_t q = (size_t)(mbar - opt.mbar_min) * pstride + d;
        if (out.I) out.I[q] = v.I;
        if (out.Q) out.Q[q] = v.Q;
        if (out.U) out.U[q] = v.U;
        if (out.r && out.r != g.r) out.r[q] = v.r;
        if (out.Ir && out.Ir != g.Ir) out.Ir[q] = v.Ir;
        if (out.g) out.g[q] = v.g;
        if (out.sinthb) out.sinthb[q] = v.sinthb;
        if (out.lp) out.lp[q] = v.lp;
        if (out.Ie) out.Ie[q] = v.Ie;
        lastN = v.Nmax;
    }
    if (out.nmax) out.nmax[d] = lastN;
}

// element-wise physics
__global__ void __launch_bounds__(EW_THREADS) ew_velocity_kernel(kgeo_models M, double a, size_t n, const double* r, double* u0,
                                                                 double* u1, double* u2, double* u3) {
    EW_INDEX
    U4 u = velocity_u_lab(M, a, r[i]);
    if (u0) u0[i] = u.u0;
    if (u1) u1[i] = u.u1;
    if (u2) u2[i] = u.u2;
    if (u3) u3[i] = u.u3;
}
__global__ void __launch_bounds__(EW_THREADS) ew_bfield_kernel(kgeo_models M, double a, size_t n, const double* r, double* B1,
                                                               double* B2, double* B3, double* bsq) {
    EW_INDEX
    double b1, b2, b3;
    bfield_lab(M, a, r[i], &b1, &b2, &b3);
    if (B1) B1[i] = b1;
    if (B2) B2[i] = b2;
    if (B3) B3[i] = b3;
    if (bsq) bsq[i] = bfield_bsq(M, a, r[i], velocity_u_lab(M, a, r[i]));
}
__global__ void __launch_bounds__(EW_THREADS) ew_bfield_omega_kernel(kgeo_models M, double a, size_t n, const double* r, double* omega) {
    EW_INDEX
    double b1, b2, b3, om;
    bfield_lab_omega(M, a, r[i], &b1, &b2, &b3, &om);
    omega[i] = om;
}
__global__ void __launch_bounds__(EW_THREADS) ew_emissivity_kernel(kgeo_models M, double a, size_t n, const double* r,
                                                                   const double* g, const double* sinthb, const double* Bmag,
                                                                   double nu_obs, double specind, double* j) {
    EW_INDEX
    j[i] = emissivity_jrest(M, a, r[i], g != nullptr, g ? g[i] : 1.0, sinthb != nullptr, sinthb ? sinthb[i] : 1.0,
                            Bmag ? Bmag[i] : 0.0, nu_obs, specind);
}

// FP64 roofline probe: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double seed, double* sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1.0e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) sink[0] = s;   // never true; keeps the chains alive
}

// RAII helpers of the host-buffer entry points
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
    template <class T>
    T* as() const { return (T*)p; }
};
// Stream-ordered scratch of the host-buffer entry points.  cudaFree of the ~4 GB a BASELINE cfg 3 call uses took up to 200 ms
// per call (measured: 305 ms to the last byte on the host, 490-515 ms until the call returned), so the scratch comes from the
// device's default memory pool, whose release threshold is raised once: freed blocks stay in the pool and the next call gets
// them back without a trip to the driver.  kgeo_trim_memory() hands the cached blocks back.
struct PoolBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~PoolBuf() {
        if (p) cudaFreeAsync(p, st);
    }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        st = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 8, stream);
    }
    template <class T>
    T* as() const { return (T*)p; }
};
static cudaError_t keep_pool_memory(int device) {
    static std::atomic<unsigned long long> done{0};
    if (device < 0 || device >= 64 || (done.load() >> device) & 1ULL) return cudaSuccess;
    cudaMemPool_t pool;
    cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
    if (e != cudaSuccess) return e;
    unsigned long long thr = ~0ULL;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    if (e == cudaSuccess) done.fetch_or(1ULL << device);
    return e;
}
// cudaSetDevice for the duration of a host-buffer entry point; the caller's current device is restored on every return path
// (a torch process would otherwise find torch.cuda.current_device() changed behind its back)
struct DeviceGuard {
    int prev = -1;
    cudaError_t set(int device) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) { prev = -1; return e; }
        return cudaSetDevice(device);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
// declared AFTER buffers a stream still reads from, so that it runs before their destructors on early (error) returns
struct SyncOnExit {
    cudaStream_t* s;
    ~SyncOnExit() {
        if (*s) cudaStreamSynchronize(*s);
    }
};
struct StreamPair {
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    ~StreamPair() {
        for (int i = 0; i < 2; ++i) {
            if (done[i]) cudaEventDestroy(done[i]);
            if (drained[i]) cudaEventDestroy(drained[i]);
        }
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
    }
};


// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

int kgeo_abi_version(void) { return KGEO_B200_ABI_VERSION; }
const char* kgeo_last_error(void) { return g_err; }
int kgeo_device_count(int* count) {
    if (!count) return fail_arg(KGEO_E_BADARG, "count is NULL");
    *count = 0;
    CU(cudaGetDeviceCount(count));
    return 0;
}
int kgeo_ray_workspace_slots(void) { return NW; }
size_t kgeo_ray_workspace_doubles(size_t npix) { return ws_doubles(npix); }
long long kgeo_launch_count(void) { return g_launches.load(); }

#define EW_ENTRY(kernel, ...)                                                             \
    if (n == 0) return 0;                                                                 \
    kernel<<<grid_for(n, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
    LAUNCHED(#kernel);                                                                    \
    return 0;

int kgeo_special_rf(size_t n, const double* x, const double* y, const double* z, double* out, void* stream) {
    if (n && (!x || !y || !z || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_rf: NULL pointer");
    EW_ENTRY(ew_rf_kernel, n, x, y, z, out)
}
int kgeo_special_rd(size_t n, const double* x, const double* y, const double* z, double* out, void* stream) {
    if (n && (!x || !y || !z || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_rd: NULL pointer");
    EW_ENTRY(ew_rd_kernel, n, x, y, z, out)
}
int kgeo_special_rj(size_t n, const double* x, const double* y, const double* z, const double* p, double* out, void* stream) {
    if (n && (!x || !y || !z || !p || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_rj: NULL pointer");
    EW_ENTRY(ew_rj_kernel, n, x, y, z, p, out)
}
int kgeo_special_rc(size_t n, const double* x, const double* y, double* out, void* stream) {
    if (n && (!x || !y || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_rc: NULL pointer");
    EW_ENTRY(ew_rc_kernel, n, x, y, out)
}
int kgeo_special_ellipk(size_t n, const double* m, double* out, void* stream) {
    if (n && (!m || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_ellipk: NULL pointer");
    EW_ENTRY(ew_K_kernel, n, m, out)
}
int kgeo_special_ellipkinc(size_t n, const double* phi, const double* m, double* out, void* stream) {
    if (n && (!phi || !m || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_ellipkinc: NULL pointer");
    EW_ENTRY(ew_F_kernel, n, phi, m, out)
}
int kgeo_special_ellipeinc(size_t n, const double* phi, const double* m, double* out, void* stream) {
    if (n && (!phi || !m || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_ellipeinc: NULL pointer");
    EW_ENTRY(ew_E_kernel, n, phi, m, out)
}
int kgeo_special_ellippi(size_t n, const double* nn, const double* phi, const double* m, double* out, void* stream) {
    if (n && (!nn || !phi || !m || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_special_ellippi: NULL pointer");
    EW_ENTRY(ew_Pi_kernel, n, nn, phi, m, out)
}
int kgeo_special_ellipj(size_t n, const double* u, const double* m, double* sn, double* cn, double* dn, double* am,
                        void* stream) {
    if (n && (!u || !m)) return fail_arg(KGEO_E_BADARG, "kgeo_special_ellipj: NULL pointer");
    EW_ENTRY(ew_jacobi_kernel, n, u, m, sn, cn, dn, am)
}

// per-ray setup = two kernels on one stream: ray_setup_kernel (roots, classes, observer-side integrals; theta blocks left
// pending) and theta_setup_kernel (the theta-function blocks); see theta_setup in kg_geodesic.cuh
static int launch_setup(const Observer& o, size_t npix, const double* alpha, const double* beta, int mode, double* w,
                         int32_t* flags, cudaStream_t stream) {
#if KG_SETUP_SPLIT
    ray_setup_kernel<true, 1><<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, stream>>>(o, npix, alpha, beta, mode, w, flags);
    LAUNCHED("ray_setup_kernel");
    ray_setup_kernel<true, 2><<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, stream>>>(o, npix, alpha, beta, mode, w, flags);
    LAUNCHED("ray_setup_kernel");
#else
    ray_setup_kernel<true, 0><<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, stream>>>(o, npix, alpha, beta, mode, w, flags);
    LAUNCHED("ray_setup_kernel");
#endif
    if (!(mode & M_CARLSON_ONLY)) {
        theta_setup_kernel<<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, stream>>>(o, npix, w, flags);
        LAUNCHED("theta_setup_kernel");
    }
    return 0;
}

int kgeo_ray_setup(const kgeo_observer* obs, size_t npix, const double* alpha, const double* beta, int mode, double* w,
                   int32_t* flags, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && (!alpha || !beta || !w || !flags)) return fail_arg(KGEO_E_BADARG, "kgeo_ray_setup: NULL pointer");
    if (npix == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    return launch_setup(o, npix, alpha, beta, mode, w, flags, (cudaStream_t)stream);
}

int kgeo_ray_setup_conserved(const kgeo_observer* obs, size_t npix, const double* lam, const double* eta, const double* s_o,
                             int mode, int phi_and_t, double* w, int32_t* flags, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && (!lam || !eta || !w || !flags)) return fail_arg(KGEO_E_BADARG, "kgeo_ray_setup_conserved: NULL pointer");
    if (phi_and_t && isinf(obs->r_o)) return fail_arg(KGEO_E_RANGE, "kgeo_ray_setup_conserved: r_o must be finite for the phi, t, sigma integrals");
    if (npix == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    cudaStream_t st = (cudaStream_t)stream;
    if (phi_and_t) ray_setup_conserved_kernel<true><<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, st>>>(o, npix, lam, eta, s_o, mode, w, flags);
    else ray_setup_conserved_kernel<false><<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, st>>>(o, npix, lam, eta, s_o, mode, w, flags);
    LAUNCHED("ray_setup_conserved_kernel");
    if (!(mode & M_CARLSON_ONLY)) {
        theta_setup_kernel<<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, st>>>(o, npix, w, flags);
        LAUNCHED("theta_setup_kernel");
    }
    return 0;
}

int kgeo_track_integrals(const kgeo_observer* obs, size_t npix, const double* w, const int32_t* flags, size_t ntau,
                         const double* tau, size_t tau_ld, int phi_and_t, double* th, double* Gph, double* Gt, double* r,
                         double* Iph, double* It, double* Isig, size_t ld, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && ntau && (!w || !flags || !tau)) return fail_arg(KGEO_E_BADARG, "kgeo_track_integrals: NULL pointer");
    if (ld < npix || tau_ld < npix) return fail_arg(KGEO_E_BADARG, "kgeo_track_integrals: leading dimension smaller than npix");
    if (npix == 0 || ntau == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    dim3 grid(grid_for(npix, TRK_THREADS), (unsigned)(ntau < 1024 ? ntau : 1024));
    track_integrals_kernel<<<grid, TRK_THREADS, 0, (cudaStream_t)stream>>>(o, npix, w, flags, (int)ntau, tau, tau_ld, phi_and_t != 0,
                                                                          IntegralsOut{th, Gph, Gt, r, Iph, It, Isig}, ld);
    LAUNCHED("track_integrals_kernel");
    return 0;
}

int kgeo_ray_slot(size_t npix, const double* w, int slot, double* out, void* stream) {
    if (npix && (!w || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_ray_slot: NULL pointer");
    if (slot < 0 || slot >= NW) return fail_arg(KGEO_E_BADARG, "kgeo_ray_slot: slot out of range");
    if (npix == 0) return 0;
    ray_slot_kernel<<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, (cudaStream_t)stream>>>(npix, w, slot, out);
    LAUNCHED("ray_slot_kernel");
    return 0;
}

int kgeo_ray_summary(const kgeo_observer* obs, size_t npix, const double* alpha, const double* beta, int mode, double* roots,
                     int32_t* rclass, double* tau_tot, double* uplus, double* uminus, double* uratio, double* a2uminus,
                     double* outside, double* thplus, double* thminus, void* stream) {
    int rc = check_observer(obs, true, true);
    if (rc) return rc;
    if (npix && (!alpha || !beta)) return fail_arg(KGEO_E_BADARG, "kgeo_ray_summary: NULL pointer");
    if (npix == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    ray_summary_kernel<<<grid_for(npix, SETUP_THREADS), SETUP_THREADS, 0, (cudaStream_t)stream>>>(
        o, npix, alpha, beta, mode, roots, rclass, tau_tot, uplus, uminus, uratio, a2uminus, outside, thplus, thminus);
    LAUNCHED("ray_summary_kernel");
    return 0;
}

int kgeo_mino_total(const kgeo_observer* obs, size_t n, const double* roots, double* out, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (n && (!roots || !out)) return fail_arg(KGEO_E_BADARG, "kgeo_mino_total: NULL pointer");
    if (n == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    mino_total_kernel<<<grid_for(n, SETUP_THREADS), SETUP_THREADS, 0, (cudaStream_t)stream>>>(o, n, roots, out);
    LAUNCHED("mino_total_kernel");
    return 0;
}

extern "C++" {
template <bool FULL, int PASS, bool HAS_TAU>
static int launch_track_pass(const Observer& o, size_t npix, const double* w, const int32_t* flags, size_t ntau, int k_begin,
                             int k_end, double maxtaufrac, const double* tau, size_t tau_ld, TrackOut out, size_t ld,
                             cudaStream_t stream) {
    constexpr int WARPS = TRK_WARPS;
    // samples [k_begin, k_end) of an ntau-sample track; output (and tau) rows are indexed by (k - k_begin)
    const size_t tiles_ = (npix + TRK_RPW * TRK_GROUPS - 1) / (TRK_RPW * TRK_GROUPS);
    if (tiles_ > 0x7fffffffu) return fail_arg(KGEO_E_BADARG, "kgeo_tracks: too many rays in one launch");
    const unsigned tiles = (unsigned)tiles_;
    const int ns = k_end - k_begin;
    // enough CTAs to fill 148 SMs several times over even for a single image row; a CTA walks whole 8-sample blocks
    const int want = 148 * 8;
    int ychunks = (int)((want + tiles - 1) / tiles);
    int per = (ns + ychunks - 1) / ychunks;
    const int blk = WARPS * (32 / TRK_RPW);           // samples one CTA-wide pass covers
    per = ((per + blk - 1) / blk) * blk;
    ychunks = (ns + per - 1) / per;
    if (ychunks > 65535) return fail_arg(KGEO_E_BADARG, "kgeo_tracks: too many Mino-time samples");
    dim3 grid(tiles, (unsigned)ychunks);
    const size_t smem = TRK_SMEM;
    CU(cudaFuncSetAttribute(track_kernel<FULL, PASS, HAS_TAU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    track_kernel<FULL, PASS, HAS_TAU><<<grid, 32 * WARPS, smem, stream>>>(o, npix, w, flags, (int)ntau, maxtaufrac, tau, tau_ld,
                                                                           out, ld, k_begin, k_end, per);
    LAUNCHED("track_kernel");
    return 0;
}
}  // extern "C++"

static int launch_hot_pass(const Observer& o, size_t npix, const double* w, const int32_t* flags, size_t ntau, int k_begin, int k_end,
                           double maxtaufrac, const double*, size_t, TrackOut out, size_t ld, cudaStream_t stream) {
    const size_t tiles_ = (npix + TRK_RPW * TRK_GROUPS - 1) / (TRK_RPW * TRK_GROUPS);
    if (tiles_ > 0x7fffffffu) return fail_arg(KGEO_E_BADARG, "kgeo_tracks: too many rays in one launch");
    const unsigned tiles = (unsigned)tiles_;
    const int ns = k_end - k_begin;
    const int want = 148 * 8;
    int ychunks = (int)((want + tiles - 1) / tiles);
    int per = (ns + ychunks - 1) / ychunks;
    const int blk = HOT_NS * HOT_WARPS * (32 / TRK_RPW);    // samples one CTA-wide pass covers
    per = ((per + blk - 1) / blk) * blk;
    ychunks = (ns + per - 1) / per;
    if (ychunks > 65535) return fail_arg(KGEO_E_BADARG, "kgeo_tracks: too many Mino-time samples");
    dim3 grid(tiles, (unsigned)ychunks);
    const size_t smem = HOT_SMEM;
    CU(cudaFuncSetAttribute(track_hot_kernel<HOT_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    track_hot_kernel<HOT_NS><<<grid, 32 * HOT_WARPS, smem, stream>>>(o, npix, w, flags, (int)ntau, maxtaufrac, out, ld, k_begin,
                                                                      k_end, per);
    LAUNCHED("track_hot_kernel");
    return 0;
}

// (Forking the fix-up pass onto a side stream so that its few long-running CTAs overlap the hot pass was measured: 20.02
// against 20.08 ms -- not worth a second stream inside the library.)
#define TRK_ARGS o, npix, w, flags, ntau, k_begin, k_end, maxtaufrac, tau, tau_ld, out, ld, stream
// full: 0 = r, theta only; 1 = all coordinates, hot + fix-up launches; 2 = all coordinates, every ray through the complete
// code in one launch (the A/B reference of the tests; also KGEO_B200_TRK_PASSES=fused)
static int launch_tracks(const Observer& o, size_t npix, const double* w, const int32_t* flags, size_t ntau, int k_begin,
                         int k_end, double maxtaufrac, const double* tau, size_t tau_ld, int full, TrackOut out, size_t ld,
                         cudaStream_t stream) {
    static const bool fused_env = [] {
        const char* e = getenv("KGEO_B200_TRK_PASSES");
        return e && !strcmp(e, "fused");
    }();
    if (tau) return full ? launch_track_pass<true, P_ALL, true>(TRK_ARGS) : launch_track_pass<false, P_ALL, true>(TRK_ARGS);
    if (!full) return launch_track_pass<false, P_ALL, false>(TRK_ARGS);
    if (full == 2 || fused_env) return launch_track_pass<true, P_ALL, false>(TRK_ARGS);
    const int rc = launch_hot_pass(TRK_ARGS);
    if (rc) return rc;
    return launch_track_pass<true, P_FIXUP, false>(TRK_ARGS);
}
#undef TRK_ARGS

int kgeo_tracks(const kgeo_observer* obs, size_t npix, const double* w, const int32_t* flags, size_t ntau, double maxtaufrac,
                const double* tau, size_t tau_ld, int full, double* t, double* r, double* th, double* ph, double* sig,
                double* tau_out, size_t ld, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && (!w || !flags)) return fail_arg(KGEO_E_BADARG, "kgeo_tracks: NULL workspace");
    if (ld < npix || (tau && tau_ld < npix)) return fail_arg(KGEO_E_BADARG, "kgeo_tracks: leading dimension smaller than npix");
    if (ntau > 0x7fffffff) return fail_arg(KGEO_E_BADARG, "kgeo_tracks: ntau too large");
    if (npix == 0 || ntau == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    TrackOut out{t, r, th, ph, sig, tau_out};
    return launch_tracks(o, npix, w, flags, ntau, 0, (int)ntau, maxtaufrac, tau, tau_ld, full, out, ld, (cudaStream_t)stream);
}

int kgeo_raytrace(const kgeo_observer* obs, size_t npix, const double* alpha, const double* beta, int mode, size_t ngeo,
                  double maxtaufrac, double* w, int32_t* flags, double* t, double* r, double* th, double* ph, double* sig,
                  double* tau_out, size_t ld, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    rc = kgeo_ray_setup(obs, npix, alpha, beta, mode, w, flags, stream);
    if (rc) return rc;
    return kgeo_tracks(obs, npix, w, flags, ngeo, maxtaufrac, nullptr, 0, 1, t, r, th, ph, sig, tau_out, ld, stream);
}

int kgeo_fieldline_guesses(const kgeo_observer* obs, size_t npix, size_t ngeo, const double* r, const double* th, const double* tau,
                           size_t ld, const double* alpha, const double* beta, const double* w, const int32_t* flags, int model,
                           double pval, double psitarget, int neqmax, double maxtaufrac, int maxlen, int32_t* count,
                           int32_t* maxcount, double* guesses, double* taumax, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && (!alpha || !beta || !w || !flags)) return fail_arg(KGEO_E_BADARG, "kgeo_fieldline_guesses: NULL pointer");
    if (npix && ngeo && (!r || !th || !tau)) return fail_arg(KGEO_E_BADARG, "kgeo_fieldline_guesses: NULL track arrays");
    if (ld < npix || ngeo > 0x7fffffff) return fail_arg(KGEO_E_BADARG, "kgeo_fieldline_guesses: bad leading dimension / ngeo");
    if (model < PSI_PARA || model > PSI_NONE) return fail_arg(KGEO_E_MODEL, "kgeo_fieldline_guesses: unknown psi model");
    if (npix == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    PsiSpec ps{model, pval, psitarget};
    cudaStream_t st = (cudaStream_t)stream;
    if (maxlen <= 0) {            // pass 0: counts per ray and their maximum
        if (!count || !maxcount) return fail_arg(KGEO_E_BADARG, "kgeo_fieldline_guesses: count / maxcount missing");
        CU(cudaMemsetAsync(maxcount, 0, sizeof(int32_t), st));
        fieldline_guess_kernel<0><<<grid_for(npix, 256), 256, 0, st>>>(o, ps, npix, (int)ngeo, r, th, tau, ld, alpha, beta, w, flags,
                                                                      neqmax, maxtaufrac, count, maxcount, 0, nullptr, taumax);
    } else {
        if (!guesses) return fail_arg(KGEO_E_BADARG, "kgeo_fieldline_guesses: guesses missing");
        fieldline_guess_kernel<1><<<grid_for(npix, 256), 256, 0, st>>>(o, ps, npix, (int)ngeo, r, th, tau, ld, alpha, beta, w, flags,
                                                                      neqmax, maxtaufrac, nullptr, nullptr, maxlen, guesses, nullptr);
    }
    LAUNCHED("fieldline_guess_kernel");
    return 0;
}

int kgeo_fieldline_solve(const kgeo_observer* obs, size_t npix, size_t nguess, const double* beta, const double* w,
                         const int32_t* flags, int model, double pval, double psitarget, const double* guesses, double tol,
                         int maxiter, double* tau_out, double* r_out, double* th_out, double* signpr, double* signpth, double* neq,
                         int32_t* iterations, int32_t* nfailed, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    const size_t n = npix * nguess;
    if (n && (!beta || !w || !flags || !guesses || !tau_out || !r_out || !th_out || !signpr || !signpth))
        return fail_arg(KGEO_E_BADARG, "kgeo_fieldline_solve: NULL pointer");
    if (model < PSI_PARA || model > PSI_NONE) return fail_arg(KGEO_E_MODEL, "kgeo_fieldline_solve: unknown psi model");
    if (iterations) *iterations = 0;
    if (nfailed) *nfailed = 0;
    if (n == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    PsiSpec ps{model, pval, psitarget};
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf dbl, ints;
    CU(dbl.alloc(5 * n * sizeof(double)));
    CU(ints.alloc((2 * n + 2) * sizeof(int32_t)));
    SecantState S;
    S.p = dbl.as<double>();
    S.p1 = S.p + n;
    S.q0 = S.p1 + n;
    S.q1 = S.q0 + n;
    S.pnew = S.q1 + n;
    S.active = ints.as<int32_t>();
    S.fail = S.active + n;
    S.gflags = S.fail + n;
    const unsigned grid = grid_for(n, 256);
    fieldline_secant_start_kernel<<<grid, 256, 0, st>>>(n, guesses, S);
    LAUNCHED("fieldline_secant_start_kernel");
    fieldline_eval_kernel<<<grid, 256, 0, st>>>(o, ps, npix, n, w, flags, guesses, S.p, S.q0, nullptr, nullptr);
    LAUNCHED("fieldline_eval_kernel");
    fieldline_eval_kernel<<<grid, 256, 0, st>>>(o, ps, npix, n, w, flags, guesses, S.p1, S.q1, nullptr, nullptr);
    LAUNCHED("fieldline_eval_kernel");
    int it = 0;
    bool converged = false;
    const double* result = S.p;
    for (; it < maxiter; ++it) {      // scipy.optimize.newton's secant loop: every element iterates until ALL have converged
        int32_t g[2] = {0, 0};
        CU(cudaMemsetAsync(S.gflags, 0, 2 * sizeof(int32_t), st));
        fieldline_secant_step_kernel<<<grid, 256, 0, st>>>(n, tol, S);
        LAUNCHED("fieldline_secant_step_kernel");
        CU(cudaMemcpyAsync(g, S.gflags, sizeof(g), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (!g[0]) {                  // all derivatives zero: p = (p1 + p) / 2 for every element, without this step's update
            fieldline_average_kernel<<<grid, 256, 0, st>>>(n, S.p, S.p1, S.pnew);
            LAUNCHED("fieldline_average_kernel");
            result = S.pnew;
            converged = true;
            ++it;
            break;
        }
        double* t = S.p;              // p <- updated p
        S.p = S.pnew;
        S.pnew = t;
        result = S.p;
        if (!g[1]) {
            converged = true;
            ++it;
            break;
        }
        t = S.p1;                     // p1, p = p, p1 ; q0 = q1 ; q1 = func(p1)
        S.p1 = S.p;
        S.p = t;
        t = S.q0;
        S.q0 = S.q1;
        S.q1 = t;
        fieldline_eval_kernel<<<grid, 256, 0, st>>>(o, ps, npix, n, w, flags, guesses, S.p1, S.q1, nullptr, nullptr);
        LAUNCHED("fieldline_eval_kernel");
        result = S.p;
    }
    if (iterations) *iterations = it;
    if (!converged && nfailed) {      // the elements scipy.optimize.newton reports as not converged: |dp| >= tol at the last step
        int32_t* h = (int32_t*)malloc(n * sizeof(int32_t));
        if (!h) return fail_arg(KGEO_E_BADARG, "kgeo_fieldline_solve: out of host memory");
        cudaError_t e = cudaMemcpyAsync(h, S.fail, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        long long bad = 0, zero_der = 0;
        if (e == cudaSuccess)
            for (size_t i = 0; i < n; ++i) {
                bad += (h[i] & 1) != 0;
                zero_der += h[i] == 3;                // ~nz_der & failures
            }
        free(h);
        if (e != cudaSuccess) return fail_cuda(e, "kgeo_fieldline_solve: failure count");
        const int32_t cnt = (int32_t)(bad > 0x7fffffffLL ? 0x7fffffffLL : bad);
        *nfailed = zero_der ? -cnt : cnt;             // negated: some of them are zero-derivative elements (scipy only warns)
    }
    fieldline_finish_kernel<<<grid, 256, 0, st>>>(o, npix, n, w, flags, beta, guesses, result, tau_out, r_out, th_out, signpr,
                                                  signpth, neq);
    LAUNCHED("fieldline_finish_kernel");
    CU(cudaStreamSynchronize(st));    // the state buffers are freed on return
    return 0;
}

int kgeo_jet_stokes(const kgeo_observer* obs, size_t npix, size_t nguess, const double* alpha, const double* beta, const double* r,
                    const double* th, const double* signpr, const double* signpth, int field, double C, double pval, int usemono,
                    const double* fr_table, int64_t fr_n, double nu_parallel, int nu_ff, const double* ff_global,
                    const double* ff_rows_dev, double gammamax, double sigma, double r0, double norm, double psitarget,
                    double specind, double* I, double* Q, double* U, double* sinthb, double* density, int32_t* nu_out_of_range,
                    void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    const size_t n = npix * nguess;
    if (n && (!alpha || !beta || !r || !th || !signpr || !signpth || !I || !Q || !U))
        return fail_arg(KGEO_E_BADARG, "kgeo_jet_stokes: NULL pointer");
    if (field != JET_MONO && field != JET_POWER) return fail_arg(KGEO_E_MODEL, "kgeo_jet_stokes: unknown jet field");
    if (field == JET_MONO && (!fr_table || fr_n < 2)) return fail_arg(KGEO_E_MODEL, "kgeo_jet_stokes: the monopole needs its f(r) table");
    if (nu_ff && (!ff_global || !ff_rows_dev)) return fail_arg(KGEO_E_BADARG, "kgeo_jet_stokes: nu_parallel = 'FF' needs the field-line constants");
    if (nu_out_of_range) *nu_out_of_range = 0;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    JetSpec J{field, C, pval, usemono, fr_table, fr_table ? fr_table + fr_n : nullptr, (long long)fr_n};
    JetBoost boost = jet_boost_numeric(nu_parallel);
    if (nu_ff) {
        boost.ff = 1;
        boost.r0 = ff_global[0];
        boost.omega = ff_global[1];
        boost.Aconst = ff_global[2];
    }
    boost.gammamax = (gammamax > 0.0) ? gammamax : 0.0;
    JetBoostRows rows{nu_ff ? ff_rows_dev : nullptr};
    DevBuf scratch;
    double* gscale = nullptr;
    int32_t* flag = nullptr;
    // scratch: [gamma n][gamma_eff n][gscale 1][minkey 1][argidx 1][flag 1]
    CU(scratch.alloc((2 * n + 4) * sizeof(double)));
    double* gam = scratch.as<double>();
    double* geff = gam + n;
    unsigned long long* minkey = reinterpret_cast<unsigned long long*>(geff + n + 1);
    unsigned long long* argidx = minkey + 1;
    flag = reinterpret_cast<int32_t*>(argidx + 1);
    CU(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    if (boost.gammamax > 0.0) {
        gscale = geff + n;
        CU(cudaMemsetAsync(minkey, 0xff, 2 * sizeof(unsigned long long), st));    // both start at the largest key / index
        jet_gamma_kernel<<<grid_for(n, 256), 256, 0, st>>>(J, obs->a, n, r, th, boost, gam, geff, minkey);
        LAUNCHED("jet_gamma_kernel");
        jet_argmin_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, geff, minkey, argidx);
        LAUNCHED("jet_argmin_kernel");
        jet_gscale_kernel<<<1, 1, 0, st>>>(gam, geff, argidx, gscale);
        LAUNCHED("jet_gscale_kernel");
    }
    jet_stokes_kernel<<<grid_for(n, 256), 256, 0, st>>>(J, obs->a, obs->th_o, npix, n, alpha, beta, r, th, signpr, signpth, boost, rows,
                                                         gscale, sigma, r0, norm, psitarget, specind, I, Q, U, sinthb, density, flag);
    LAUNCHED("jet_stokes_kernel");
    if (nu_out_of_range) CU(cudaMemcpyAsync(nu_out_of_range, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));          // the scratch is freed on return
    return 0;
}

int kgeo_jet_sort(size_t npix, size_t nguess, int neqmax, const double* I, const double* Q, const double* U, const double* neq,
                  double* out_I, double* out_Q, double* out_U, void* stream) {
    if (npix && nguess && (!I || !Q || !U || !neq || !out_I || !out_Q || !out_U)) return fail_arg(KGEO_E_BADARG, "kgeo_jet_sort: NULL pointer");
    if (neqmax < 0 || nguess > 0x7fffffff) return fail_arg(KGEO_E_BADARG, "kgeo_jet_sort: bad neqmax / nguess");
    if (npix == 0) return 0;
    jet_sort_kernel<<<grid_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(npix, (int)nguess, neqmax, I, Q, U, neq, out_I, out_Q, out_U);
    LAUNCHED("jet_sort_kernel");
    return 0;
}

int kgeo_crossing_counts(const kgeo_observer* obs, size_t npix, const double* beta, const double* w, const int32_t* flags,
                         size_t ntau, const double* tau, double* neq, double* npol, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && ntau && (!beta || !w || !flags || !tau)) return fail_arg(KGEO_E_BADARG, "kgeo_crossing_counts: NULL pointer");
    if (npix == 0 || ntau == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    dim3 grid(grid_for(npix, TRK_THREADS), (unsigned)(ntau < 1024 ? ntau : 1024));
    crossing_counts_kernel<<<grid, TRK_THREADS, 0, (cudaStream_t)stream>>>(o, npix, beta, w, flags, (int)ntau, tau, neq, npol);
    LAUNCHED("crossing_counts_kernel");
    return 0;
}

int kgeo_r_equatorial(const kgeo_observer* obs, size_t npix, const double* alpha, const double* beta, int mbar_lo, int mbar_hi,
                      double* rs, double* Ir, double* Nmax, double* Imax, double* npol, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && (!alpha || !beta)) return fail_arg(KGEO_E_BADARG, "kgeo_r_equatorial: NULL pointer");
    if (mbar_lo < 0 || mbar_hi < mbar_lo) return fail_arg(KGEO_E_BADARG, "kgeo_r_equatorial: bad mbar range");
    if (npix == 0) return 0;
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    r_equatorial_kernel<<<grid_for(npix, IMG_THREADS), IMG_THREADS, 0, (cudaStream_t)stream>>>(o, npix, alpha, beta, mbar_lo,
                                                                                              mbar_hi, rs, Ir, Nmax, Imax, npol, nullptr);
    LAUNCHED("r_equatorial_kernel");
    return 0;
}

int kgeo_rho_of_req(const kgeo_observer* obs, size_t n, const double* varphi, double req, int mbar, double rho_min, double rho_max,
                    double xtol, double rtol, int maxiter, double* rho, double* alpha, double* beta, int32_t* status,
                    void* stream) {
    if (!obs) return fail_arg(KGEO_E_BADARG, "kgeo_rho_of_req: obs is NULL");
    kgeo_observer oinf = *obs;
    oinf.r_o = KG_INF;                                       // R0 = inf (kgeo/equatorial_lensing.py:19)
    int rc = check_observer(&oinf, true);
    if (rc) return rc;
    if (n && !varphi) return fail_arg(KGEO_E_BADARG, "kgeo_rho_of_req: NULL pointer");
    if (mbar < 0 || maxiter < 0 || !(xtol > 0.0)) return fail_arg(KGEO_E_BADARG, "kgeo_rho_of_req: bad mbar / tolerance");
    if (n == 0) return 0;
    Observer o = make_observer(oinf.a, oinf.r_o, oinf.th_o);
    rho_of_req_kernel<<<grid_for(n, 32), 32, 0, (cudaStream_t)stream>>>(o, n, varphi, req, mbar, rho_min, rho_max, xtol, rtol,
                                                                       maxiter, rho, alpha, beta, status);
    LAUNCHED("rho_of_req_kernel");
    return 0;
}

int kgeo_rho_library(size_t nconf, const double* spin, const double* th0, const double* req, int mbar_lo, int mbar_hi, size_t nphi,
                     const double* varphi, double xtol, double rtol, int maxiter, double* rho, int32_t* status, void* stream) {
    if (nconf && nphi && (!spin || !th0 || !req || !varphi || !rho)) return fail_arg(KGEO_E_BADARG, "kgeo_rho_library: NULL pointer");
    if (mbar_lo < 0 || mbar_hi < mbar_lo || maxiter < 0 || !(xtol > 0.0)) return fail_arg(KGEO_E_BADARG, "kgeo_rho_library: bad mbar range / tolerance");
    if (nconf == 0 || nphi == 0) return 0;
    const int norder = mbar_hi - mbar_lo + 1;
    rho_library_kernel<<<grid_for(nconf * norder * nphi, 32), 32, 0, (cudaStream_t)stream>>>(nconf, spin, th0, req, mbar_lo, norder, nphi,
                                                                                            varphi, xtol, rtol, maxiter, rho, status);
    LAUNCHED("rho_library_kernel");
    return 0;
}

int kgeo_critical_curve(double a, double r_minus, double r_plus, double sin_th0, double cos_th0, double tan_th0, size_t n,
                        double* alpha, double* beta, void* stream) {
    if (n && (!alpha || !beta)) return fail_arg(KGEO_E_BADARG, "kgeo_critical_curve: NULL pointer");
    if (n == 0) return 0;
    critical_curve_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(a, r_minus, r_plus, sin_th0, cos_th0, tan_th0, n,
                                                                             alpha, beta);
    LAUNCHED("critical_curve_kernel");
    return 0;
}

// geometry into the caller's r / Ir / imax / npol planes where they are plain local planes, into pool scratch otherwise
// (sharded: the destination may be peer memory, which the emission kernel must not read back), then the emission
static int launch_image(const kgeo_observer* obs, const kgeo_models* models, const kgeo_image_opts* opts, size_t npix,
                        const double* alpha, const double* beta, const kgeo_shard_map* map, const ImageOut& out, cudaStream_t st) {
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    const size_t norder = (size_t)(opts->mbar_max - opts->mbar_min + 1), plane = norder * npix * sizeof(double);
    int dev = 0;
    CU(cudaGetDevice(&dev));
    CU(keep_pool_memory(dev));
    PoolBuf s_r, s_Ir, s_N, s_tau, s_npol, s_out;
    const bool local = (map == nullptr);
    double *g_r = local ? out.r : nullptr, *g_Ir = local ? out.Ir : nullptr, *g_tau = local ? out.imax : nullptr,
           *g_npol = local ? out.npol : nullptr;
    if (!g_r) { CU(s_r.alloc(plane, st)); g_r = s_r.as<double>(); }
    if (!g_Ir) { CU(s_Ir.alloc(plane, st)); g_Ir = s_Ir.as<double>(); }
    if (!g_tau) { CU(s_tau.alloc(npix * sizeof(double), st)); g_tau = s_tau.as<double>(); }
    if (!g_npol && out.npol) { CU(s_npol.alloc(npix * sizeof(double), st)); g_npol = s_npol.as<double>(); }
    CU(s_N.alloc(plane, st));
    CU(s_out.alloc(npix, st));
    const unsigned grid = grid_for(npix, EMIS_THREADS);
    r_equatorial_kernel<<<grid_for(npix, IMG_THREADS), IMG_THREADS, 0, st>>>(o, npix, alpha, beta, opts->mbar_min, opts->mbar_max, g_r, g_Ir,
                                                                             s_N.as<double>(), g_tau, g_npol, s_out.as<unsigned char>());
    LAUNCHED("r_equatorial_kernel");
    const GeomPlanes g{g_r, g_Ir, s_N.as<double>(), g_tau, g_npol, s_out.as<unsigned char>()};
    if (map) {
        if (opts->polarization) image_emission_kernel<true, true><<<grid, EMIS_THREADS, 0, st>>>(o, *models, *opts, npix, alpha, beta, g, out, *map);
        else image_emission_kernel<false, true><<<grid, EMIS_THREADS, 0, st>>>(o, *models, *opts, npix, alpha, beta, g, out, *map);
    } else {
        const kgeo_shard_map id{};
        if (opts->polarization) image_emission_kernel<true, false><<<grid, EMIS_THREADS, 0, st>>>(o, *models, *opts, npix, alpha, beta, g, out, id);
        else image_emission_kernel<false, false><<<grid, EMIS_THREADS, 0, st>>>(o, *models, *opts, npix, alpha, beta, g, out, id);
    }
    LAUNCHED("image_emission_kernel");
    return 0;
}

int kgeo_equatorial_image(const kgeo_observer* obs, const kgeo_models* models, const kgeo_image_opts* opts, size_t npix,
                          const double* alpha, const double* beta, double* I, double* Q, double* U, double* r, double* Ir,
                          double* g, double* sinthb, double* nmax, double* npol, double* imax, double* lp, double* Ie,
                          void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    rc = check_models(models);
    if (rc) return rc;
    if (!opts) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image: opts is NULL");
    if (opts->mbar_min < 0 || opts->mbar_max < opts->mbar_min) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image: bad mbar range");
    if (npix && (!alpha || !beta)) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image: NULL pointer");
    if (npix == 0) return 0;
    return launch_image(obs, models, opts, npix, alpha, beta, nullptr,
                        ImageOut{I, Q, U, r, Ir, g, sinthb, nmax, npol, imax, opts->pathlength ? lp : nullptr, opts->pathlength ? Ie : nullptr},
                        (cudaStream_t)stream);
}

int kgeo_equatorial_image_sharded(const kgeo_observer* obs, const kgeo_models* models, const kgeo_image_opts* opts, size_t npix,
                                  const double* alpha, const double* beta, const kgeo_shard_map* map, double* I, double* Q,
                                  double* U, double* r, double* Ir, double* g, double* sinthb, double* nmax, double* npol,
                                  double* imax, double* lp, double* Ie, void* stream) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    rc = check_models(models);
    if (rc) return rc;
    if (!opts || !map) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_sharded: opts / map is NULL");
    if (opts->mbar_min < 0 || opts->mbar_max < opts->mbar_min) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_sharded: bad mbar range");
    if (npix && (!alpha || !beta)) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_sharded: NULL pointer");
    if (npix == 0) return 0;
    if (map->cols == 0 || map->row_step == 0 || npix % map->cols != 0)
        return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_sharded: the shard must be whole rows of cols > 0 pixels, row_step > 0");
    const uint64_t last_row = (npix / map->cols - 1) * map->row_step + map->row_offset;
    if (opts->mbar_max > opts->mbar_min && map->plane_stride < (last_row + 1) * map->cols)
        return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_sharded: plane_stride is smaller than the destination plane");
    return launch_image(obs, models, opts, npix, alpha, beta, map,
                        ImageOut{I, Q, U, r, Ir, g, sinthb, nmax, npol, imax, opts->pathlength ? lp : nullptr, opts->pathlength ? Ie : nullptr},
                        (cudaStream_t)stream);
}

// The same placement by the copy engines: one strided 3-D copy (rows x planes), no SM involved.
int kgeo_shard_scatter(const double* src, size_t nplanes, size_t npix, const kgeo_shard_map* map, double* dst, void* stream) {
    if (!map || (nplanes && npix && (!src || !dst))) return fail_arg(KGEO_E_BADARG, "kgeo_shard_scatter: NULL pointer");
    if (nplanes == 0 || npix == 0) return 0;
    if (map->cols == 0 || map->row_step == 0 || npix % map->cols != 0)
        return fail_arg(KGEO_E_BADARG, "kgeo_shard_scatter: the shard must be whole rows of cols > 0 pixels, row_step > 0");
    const size_t rows = npix / map->cols, row_bytes = map->cols * sizeof(double);
    const size_t dpitch = map->row_step * row_bytes;
    if (nplanes > 1 && (map->plane_stride * sizeof(double)) % dpitch != 0) {
        // planes that are not a whole number of destination pitches apart: one 2-D copy per plane
        for (size_t p = 0; p < nplanes; ++p)
            CU(cudaMemcpy2DAsync(dst + p * map->plane_stride + map->row_offset * map->cols, dpitch, src + p * npix, row_bytes,
                                 row_bytes, rows, cudaMemcpyDefault, (cudaStream_t)stream));
        return 0;
    }
    cudaMemcpy3DParms c = {};
    c.srcPtr = make_cudaPitchedPtr((void*)src, row_bytes, row_bytes, rows);
    c.dstPtr = make_cudaPitchedPtr((void*)(dst + map->row_offset * map->cols), dpitch, dpitch,
                                   nplanes > 1 ? map->plane_stride * sizeof(double) / dpitch : rows);
    c.extent = make_cudaExtent(row_bytes, rows, nplanes);
    c.kind = cudaMemcpyDefault;
    CU(cudaMemcpy3DAsync(&c, (cudaStream_t)stream));
    return 0;
}

// ---- peer buffers: device memory that other processes of the box map and write with plain stores ----
int kgeo_peer_alloc(size_t bytes, void** dev_ptr, kgeo_peer_handle* handle) {
    if (!dev_ptr || !handle || bytes == 0) return fail_arg(KGEO_E_BADARG, "kgeo_peer_alloc: bad argument");
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));                       // cudaMalloc, not a pool: legacy IPC handles export whole cudaMalloc allocations
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        CU(e);
    }
    static_assert(sizeof(h) == sizeof(handle->bytes), "kgeo_peer_handle must hold a cudaIpcMemHandle_t");
    memcpy(handle->bytes, &h, sizeof(h));
    *dev_ptr = p;
    return 0;
}
int kgeo_peer_open(const kgeo_peer_handle* handle, void** dev_ptr) {
    if (!dev_ptr || !handle) return fail_arg(KGEO_E_BADARG, "kgeo_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof(h));
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));   // maps the exporter's memory into the CURRENT device's address space
    *dev_ptr = p;
    return 0;
}
int kgeo_peer_close(void* dev_ptr) {
    if (!dev_ptr) return 0;
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}
int kgeo_peer_free(void* dev_ptr) {
    if (!dev_ptr) return 0;
    CU(cudaFree(dev_ptr));
    return 0;
}

int kgeo_velocity_u_lab(const kgeo_models* models, double a, size_t n, const double* r, double* u0, double* u1, double* u2,
                        double* u3, void* stream) {
    int rc = check_models(models);
    if (rc) return rc;
    if (n && !r) return fail_arg(KGEO_E_BADARG, "kgeo_velocity_u_lab: NULL pointer");
    EW_ENTRY(ew_velocity_kernel, *models, a, n, r, u0, u1, u2, u3)
}
int kgeo_bfield_lab(const kgeo_models* models, double a, size_t n, const double* r, double* B1, double* B2, double* B3,
                    double* bsq, void* stream) {
    int rc = check_models(models);
    if (rc) return rc;
    if (n && !r) return fail_arg(KGEO_E_BADARG, "kgeo_bfield_lab: NULL pointer");
    EW_ENTRY(ew_bfield_kernel, *models, a, n, r, B1, B2, B3, bsq)
}
int kgeo_bfield_omega(const kgeo_models* models, double a, size_t n, const double* r, double* omega, void* stream) {
    int rc = check_models(models);
    if (rc) return rc;
    if (n && (!r || !omega)) return fail_arg(KGEO_E_BADARG, "kgeo_bfield_omega: NULL pointer");
    EW_ENTRY(ew_bfield_omega_kernel, *models, a, n, r, omega)
}
int kgeo_emissivity_jrest(const kgeo_models* models, double a, size_t n, const double* r, const double* g, const double* sinthb,
                          const double* Bmag, double nu_obs, double specind, double* j, void* stream) {
    int rc = check_models(models);
    if (rc) return rc;
    if (n && (!r || !j)) return fail_arg(KGEO_E_BADARG, "kgeo_emissivity_jrest: NULL pointer");
    EW_ENTRY(ew_emissivity_kernel, *models, a, n, r, g, sinthb, Bmag, nu_obs, specind, j)
}

int kgeo_fp64_peak_probe(int iters, double* flops_per_launch, float* ms, void* stream) {
    if (iters <= 0 || !flops_per_launch || !ms) return fail_arg(KGEO_E_BADARG, "kgeo_fp64_peak_probe: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    CU(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(iters, 1.0, sink);   // warm-up
    LAUNCHED("dfma_probe_kernel");
    CU(cudaEventRecord(e0, st));
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    LAUNCHED("dfma_probe_kernel");
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(ms, e0, e1));
    *flops_per_launch = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// host-buffer entry points (the end-to-end path)
// ---------------------------------------------------------------------------------------------------------------------
// Sample slabs are computed on one stream while the previous slab drains to the host on another (two device slab
// buffers per output).  The host arrays may be pageable or pinned; pinned arrays get true asynchronous DMA.
int kgeo_raytrace_host(const kgeo_observer* obs, size_t npix, const double* alpha, const double* beta, int mode, size_t ngeo,
                       double maxtaufrac, double* t, double* r, double* th, double* ph, double* sig, double* tau_out,
                       int device) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    if (npix && (!alpha || !beta)) return fail_arg(KGEO_E_BADARG, "kgeo_raytrace_host: NULL pointer");
    if (ngeo > 0x7fffffff) return fail_arg(KGEO_E_BADARG, "kgeo_raytrace_host: ngeo too large");
    if (npix == 0 || ngeo == 0) return 0;
    DeviceGuard dg;
    CU(dg.set(device));
    // KGEO_B200_TRACE_HOST=1 (developer): wall-clock phases of this call on stderr
    static const bool trace = getenv("KGEO_B200_TRACE_HOST") != nullptr;
    const auto tc0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (trace)
            fprintf(stderr, "kgeo_raytrace_host: %-28s %8.2f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
    };
    Observer o = make_observer(obs->a, obs->r_o, obs->th_o);
    double* host[6] = {t, r, th, ph, sig, tau_out};
    int nout = 0;
    for (int i = 0; i < 6; ++i) nout += host[i] != nullptr;
    CU(keep_pool_memory(device));
    StreamPair sp;
    CU(cudaStreamCreateWithFlags(&sp.compute, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&sp.copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&sp.done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sp.drained[i], cudaEventDisableTiming));
    }
    PoolBuf d_alpha, d_beta, d_w, d_flags;            // freed (stream-ordered, on sp.compute) before sp goes out of scope
    CU(d_alpha.alloc(npix * sizeof(double), sp.compute));
    CU(d_beta.alloc(npix * sizeof(double), sp.compute));
    CU(d_w.alloc(ws_doubles(npix) * sizeof(double), sp.compute));
    CU(d_flags.alloc(npix * sizeof(int32_t), sp.compute));
    CU(cudaMemcpyAsync(d_alpha.p, alpha, npix * sizeof(double), cudaMemcpyHostToDevice, sp.compute));
    CU(cudaMemcpyAsync(d_beta.p, beta, npix * sizeof(double), cudaMemcpyHostToDevice, sp.compute));
    lap("streams, workspace alloc");
    rc = launch_setup(o, npix, d_alpha.as<double>(), d_beta.as<double>(), mode, d_w.as<double>(), d_flags.as<int32_t>(), sp.compute);
    if (rc) return rc;
    if (nout == 0) {
        CU(cudaStreamSynchronize(sp.compute));
        return 0;
    }
    // slab = as many samples as fit ~256 MiB per output array (whole 32-sample blocks)
    size_t slab = (size_t)(256u << 20) / (npix * sizeof(double));
    slab = (slab < 32) ? 32 : (slab / 32) * 32;
    if (slab > ngeo) slab = ngeo;
    PoolBuf d_out[2][6];
    SyncOnExit drain{&sp.copy};                       // an error return must not free slabs the copy stream still reads
    for (int b = 0; b < 2; ++b)
        for (int i = 0; i < 6; ++i)
            if (host[i]) CU(d_out[b][i].alloc(slab * npix * sizeof(double), sp.compute));
    lap("slab alloc");
    int nslab = (int)((ngeo + slab - 1) / slab);
    for (int s = 0; s < nslab; ++s) {
        int b = s & 1;
        size_t k0 = (size_t)s * slab, k1 = k0 + slab < ngeo ? k0 + slab : ngeo;
        if (s >= 2) CU(cudaStreamWaitEvent(sp.compute, sp.drained[b], 0));
        TrackOut out{host[0] ? d_out[b][0].as<double>() : nullptr, host[1] ? d_out[b][1].as<double>() : nullptr,
                     host[2] ? d_out[b][2].as<double>() : nullptr, host[3] ? d_out[b][3].as<double>() : nullptr,
                     host[4] ? d_out[b][4].as<double>() : nullptr, host[5] ? d_out[b][5].as<double>() : nullptr};
        rc = launch_tracks(o, npix, d_w.as<double>(), d_flags.as<int32_t>(), ngeo, (int)k0, (int)k1, maxtaufrac, nullptr, 0, 1,
                           out, npix, sp.compute);
        if (rc) return rc;
        CU(cudaEventRecord(sp.done[b], sp.compute));
        CU(cudaStreamWaitEvent(sp.copy, sp.done[b], 0));
        for (int i = 0; i < 6; ++i)
            if (host[i])
                CU(cudaMemcpyAsync(host[i] + k0 * npix, d_out[b][i].p, (k1 - k0) * npix * sizeof(double), cudaMemcpyDeviceToHost,
                                   sp.copy));
        CU(cudaEventRecord(sp.drained[b], sp.copy));
    }
    lap("all slabs enqueued");
    CU(cudaStreamSynchronize(sp.copy));
    CU(cudaStreamSynchronize(sp.compute));
    lap("copies drained");
    return 0;
}

// The host-buffer entry points raise the release threshold of the device's DEFAULT memory pool (keep_pool_memory): a
// process-wide setting, i.e. other cudaMallocAsync users of that device also keep their freed blocks cached until this call.
int kgeo_trim_memory(int device) {
    cudaMemPool_t pool;
    DeviceGuard dg;
    CU(dg.set(device));
    CU(cudaDeviceSynchronize());
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    CU(cudaMemPoolTrimTo(pool, 0));
    return 0;
}

int kgeo_equatorial_image_host(const kgeo_observer* obs, const kgeo_models* models, const kgeo_image_opts* opts, size_t npix,
                               const double* alpha, const double* beta, double* I, double* Q, double* U, double* r, double* Ir,
                               double* g, double* sinthb, double* nmax, double* npol, double* imax, double* lp, double* Ie,
                               int device) {
    int rc = check_observer(obs, true);
    if (rc) return rc;
    rc = check_models(models);
    if (rc) return rc;
    if (!opts) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_host: opts is NULL");
    if (opts->mbar_min < 0 || opts->mbar_max < opts->mbar_min) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_host: bad mbar range");
    if (npix && (!alpha || !beta)) return fail_arg(KGEO_E_BADARG, "kgeo_equatorial_image_host: NULL pointer");
    if (npix == 0) return 0;
    DeviceGuard dg;
    CU(dg.set(device));
    cudaStream_t st;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct Guard {
        cudaStream_t s;
        ~Guard() { cudaStreamDestroy(s); }
    } guard{st};
    const size_t norder = (size_t)(opts->mbar_max - opts->mbar_min + 1);
    constexpr int NOUT = 12;
    double* host[NOUT] = {I, Q, U, r, Ir, g, sinthb, nmax, npol, imax, opts->pathlength ? lp : nullptr,
                          opts->pathlength ? Ie : nullptr};
    const size_t planes[NOUT] = {norder, norder, norder, norder, norder, norder, norder, 1, 1, 1, norder, norder};
    CU(keep_pool_memory(device));
    PoolBuf d_alpha, d_beta, d_out[NOUT];             // pool-backed scratch (see PoolBuf), freed stream-ordered before `guard`
    CU(d_alpha.alloc(npix * sizeof(double), st));
    CU(d_beta.alloc(npix * sizeof(double), st));
    for (int i = 0; i < NOUT; ++i)
        if (host[i]) CU(d_out[i].alloc(planes[i] * npix * sizeof(double), st));
    CU(cudaMemcpyAsync(d_alpha.p, alpha, npix * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_beta.p, beta, npix * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = kgeo_equatorial_image(obs, models, opts, npix, d_alpha.as<double>(), d_beta.as<double>(), d_out[0].as<double>(),
                               d_out[1].as<double>(), d_out[2].as<double>(), d_out[3].as<double>(), d_out[4].as<double>(),
                               d_out[5].as<double>(), d_out[6].as<double>(), d_out[7].as<double>(), d_out[8].as<double>(),
                               d_out[9].as<double>(), d_out[10].as<double>(), d_out[11].as<double>(), st);
    if (rc) return rc;
    for (int i = 0; i < NOUT; ++i)
        if (host[i]) CU(cudaMemcpyAsync(host[i], d_out[i].p, planes[i] * npix * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
