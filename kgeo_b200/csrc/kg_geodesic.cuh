// Per-ray setup and per-(ray, Mino time) evaluation of the analytic Kerr null geodesic (Gralla & Lupsasca 2019a,b).
//
// Replaces, for the hot path, kgeo/kerr_raytracing_utils.py (uplus_uminus :276-328, radial_roots :330-380,
// mino_total :382-548, n_equatorial_crossings :605-684, n_poloidal_orbits :551-603, is_outside_crit :794-826) and
// kgeo/kerr_raytracing_ana.py (th_integrate :195-365, r_integrate :368-633, S1_S2 :638-663, R1_R2 :666-683).
//
// Layout: `ray_setup` reduces a pixel (alpha, beta) to NW doubles + one flag word ("ray constants"): everything that
// does not depend on the Mino time, including the observer-side antiderivatives and the complete integrals needed
// for period bookkeeping.  `theta_eval` / `radial_eval` then cost one Jacobi ladder and one shared Carlson bundle
// each per Mino-time sample.  The constants are addressed through an accessor so the same code runs from a
// register array (fused setup), shared memory, or a strided SoA workspace in HBM.
//
// What is deliberately different from the reference's arithmetic (same mathematics):
//  * the quartic is solved in real arithmetic (resolvent cubic + two quadratics), branch choices as in the
//    reference's complex solve;
//  * amplitudes are never formed: F(am(u)|m) = u, sin/cos of the amplitude are sn/cn, period reduction is done
//    on u with the K(m) that falls out of the AGM ladder;
//  * E - F is evaluated as -(m/3) s^3 RD directly (no cancellation, no 1/k, so the a -> 0 limit needs no
//    special case), and all Pi's of one amplitude share one duplication track.
#pragma once
#include "kg_special.cuh"
#include "kg_theta_layout.cuh"

namespace kg {

constexpr double kMinSpin = 1.0e-4;  // kgeo/kerr_raytracing_utils.py:9
constexpr double kEP = 1.0e-12;      // kgeo/kerr_raytracing_utils.py:12

// ------------------------------------------------------------------------------------------------------------
struct Observer {
    double a, r_o, th_o;
    double a2, rplus, rminus, sth, cth, cth2;
    double c_iph, c_it;   // 2a/(r+ - r-), 4/(r+ - r-) of the I_phi / I_t assembly (kerr_raytracing_ana.py:627-630)
    int ro_inf, lowspin, north;
};

KG_HD Observer make_observer(double a, double r_o, double th_o) {
    Observer o;
    o.a = a;
    o.r_o = r_o;
    o.th_o = th_o;
    o.a2 = a * a;
    double q = sqrt(1.0 - a * a);
    o.rplus = 1.0 + q;
    o.rminus = 1.0 - q;
    o.c_iph = 2.0 * a / (o.rplus - o.rminus);
    o.c_it = 4.0 / (o.rplus - o.rminus);
    o.sth = sin(th_o);
    o.cth = cos(th_o);
    o.cth2 = o.cth * o.cth;
    o.ro_inf = isinf(r_o) ? 1 : 0;
    o.lowspin = (fabs(a) < kMinSpin) ? 1 : 0;
    o.north = (th_o < kHalfPi) ? 1 : 0;
    return o;
}

// ---- ray-constant slots --------------------------------------------------------------------------------------
enum Slot {
    S_LAM = 0, S_ETA, S_TAUTOT,
    // polar motion
    T_W,      // sqrt(|a^2 u_-|): snarg = T_X0 - T_W * tau
    T_X0,     // T_W * s_o * Gth_o
    T_M,      // reciprocal-modulus parameter m = k/(k-1) in [0,1)
    T_SQ,     // sqrt(1-m)
    T_K,      // K(m)
    T_KPAR,   // Legendre parameter k < 0 of the reference's formulas
    T_A, T_B, // theta = acos(T_A*sinPhi) (ordinary) | acos(h*sqrt(T_A + T_B sin^2 nu)) (vortical, h in flags)
    T_N,      // characteristic of the Pi in G_phi
    T_D2, T_D3, T_GPH0,         // G_phi = D3*(snarg + pi_tail) + D2*j - GPH0
    T_C1, T_C2, T_C3, T_GT0,    // G_t   = C3*snarg + C2*j + C1 * s^3 RD     - GT0
    T_F0, T_KC,                 // F0 = F(asin(cos th_o/sqrt u+)|k), K(k): crossing bookkeeping (ordinary only)
    // radial motion, common
    R_K,      // parameter k2 | k3 | k4
    R_SC,     // X = R_X0 - R_SC * tau
    R_X0,
    R_KC,     // K(k)  (complete), R_EC = E(k)
    R_EC,
    R_N0, R_N1, R_N2,           // three Pi characteristics
    R_PC0, R_PC1, R_PC2,        // Pi(n_i|k) - K(k)
    R_PO0, R_PO1, R_PO2,        // observer-side values: case 2: Pi(n_i;aux|k); case 3: R1(al_i,aux); case 4: S1(al_i,aux)
    R_EO,                       // case 2: E(aux|k);  case 3: R2(al0,aux);  case 4: S2(g0,aux)
    R_Q0, R_Q1, R_Q2, R_Q3, R_Q4, R_Q5, R_Q6, R_Q7, R_Q8, R_Q9, R_Q10,
    T_ISQ,                      // 1 / T_SQ
    T_I2K, R_I2K,               // 1/(2 K(m)), 1/(2 K(k)) (0 when K is infinite)
    // per-ray reciprocals / prefactors of the case-3 and case-4 algebra (see radial_eval): no per-sample division by a constant
    R_C0, R_C1, R_C2, R_C3, R_C4, R_C5, R_C6, R_C7, R_C8, R_C9, R_C10, R_C11, R_C12,
    R_RDC,                      // RD(0, 1-k, 1) of the radial parameter (kept for the theta-block setup)
    // theta-function path (kg_theta.cuh): polar block = base + 1 Pi, radial block = base + 3 Pi
    T_TH_G, T_TH_IM, T_TH_SLOPE, T_TH_COEF,       // 1 + (E/K - 1)/m, 1/m, G_phi slope and coefficient
    T_TH, T_TH_END = T_TH + TB_SIZE + TP_SIZE - 1,
    R_TH, R_TH_END = R_TH + TB_SIZE + 3 * TP_SIZE - 1,
    // AGM ladders (Carlson fallback only) LAST: the hot pass copies slots [0, T_LAD) of a ray group with one bulk copy
    T_LAD, T_LAD_END = T_LAD + kLadderSize - 1,   // AGM ladder of the polar Jacobi parameter m
    R_LAD, R_LAD_END = R_LAD + kLadderSize - 1,   // AGM ladder of the radial parameter k
    NW
};
// ray_setup<PHI_T = false, LADDER = false> (image / lensing kernels: register- or local-memory workspace) touches only the
// slots below R_C0; those kernels size their per-thread workspace with NW_LITE instead of NW
constexpr int NW_LITE = R_C0;
constexpr int NW_HOT = T_LAD;       // slots the theta-function-only hot pass reads

// ---- layout of the ray-constant workspace in HBM ----------------------------------------------------------------------------
// Rays are stored in GROUPS OF FOUR (the track kernels' unit of work: four consecutive pixels = one 32-byte sector of every
// output row): group g owns the contiguous block w[g * NW * 4 ... ), laid out [slot][ray in group].  A track CTA therefore
// stages a group's constants with one contiguous copy (the hot pass: ONE cp.async.bulk of NW_HOT * 32 bytes), and the setup
// kernels write whole 32-byte sectors.  Buffers hold ws_doubles(npix) doubles (npix rounded up to a multiple of four rays).
constexpr int kRaysPerGroup = 4;
constexpr int kGroupDoubles = NW * kRaysPerGroup;
KG_HD size_t ws_offset(size_t ray) { return (ray / kRaysPerGroup) * (size_t)kGroupDoubles + (ray % kRaysPerGroup); }
KG_HD size_t ws_doubles(size_t npix) { return ((npix + kRaysPerGroup - 1) / kRaysPerGroup) * (size_t)kGroupDoubles; }
// case-specific meaning of the R_Q* slots
//  case 2: Q0..Q3 = r1..r4, Q4 = sqrt(r31 r42), Q5 = H_o
//  case 3: Q0 = r1, Q1 = r2, Q2 = A, Q3 = B, Q4..Q6 = al0, alp, alm, Q7..Q9 = p1(al_i), Q10 = sqrt(AB)
//  case 4: Q0 = a2, Q1 = b1, Q2 = g0, Q3 = gp, Q4 = gm, Q5..Q7 = p2(al_i), Q8 = pref, Q9 = x4rp, Q10 = x4rm

enum Flag {
    F_TH_ORD = 1, F_TH_VORT = 2,          // polar class (neither: theta outputs stay 0, as in the reference)
    F_CASE2 = 4, F_CASE3 = 8, F_CASE4 = 16,
    F_SPOS = 32,                          // s_o = my_sign(beta) = +1
    F_HPOS = 64,                          // h = +1 (th_o <= pi/2)
    F_REGION_SHIFT = 8,                   // bits 8..10: region label 0..4 of radial_roots
    F_OUTSIDE = 1 << 12,                  // is_outside_crit
    F_LOWSPIN = 1 << 13,
    F_TH_THETA = 1 << 14,                 // polar motion may use the theta-function path
    F_R_THETA = 1 << 15,                  // radial motion may use the theta-function path
    F_R_K100 = 1 << 16,                   // ... and its three Pi blocks have kinds (1, 0, 0): rays that turn outside the horizon
    F_R_K111 = 1 << 17,                   // ... kinds (1, 1, 1): plunging rays (see theta_eval_motion)
    F_TH_PENDING = 1 << 18,               // theta blocks still to be built by theta_setup (two-kernel setup)
    F_R_PENDING = 1 << 19
};

struct Roots {
    double re[4], im[4];   // r1..r4 (im of r1 = -im of r2, im of r3 = -im of r4)
    int c12, c34;          // complex pair flags
    int region;
};

// Individually rounded IEEE operations (no FMA contraction).  The few quantities that decide a CLAMP or a CLASS in the
// reference (eta, u_+, u_-) are formed with exactly NumPy's operation order: a one-ulp difference in u_+ next to the
// clamp u_+ >= cos^2(th_o) moves the observer's polar amplitude by sqrt(ulp) ~ 1e-8 (rays with beta = 0 start exactly at a
// polar turning point), which would show up as a 1e-8 shift of theta(tau) and phi(tau) against the reference.
#if defined(__CUDA_ARCH__)
KG_HD double mul_rn(double x, double y) { return __dmul_rn(x, y); }
KG_HD double add_rn(double x, double y) { return __dadd_rn(x, y); }
KG_HD double div_rn(double x, double y) { return __ddiv_rn(x, y); }
KG_HD double sqrt_rn(double x) { return __dsqrt_rn(x); }
#else
KG_HD double mul_rn(double x, double y) { volatile double r = x * y; return r; }
KG_HD double add_rn(double x, double y) { volatile double r = x + y; return r; }
KG_HD double div_rn(double x, double y) { volatile double r = x / y; return r; }
KG_HD double sqrt_rn(double x) { volatile double r = sqrt(x); return r; }
#endif

// eta = (alpha^2 - a^2) cos^2(th_o) + beta^2 (kgeo/kerr_raytracing_ana.py:77)
KG_HD double conserved_eta(const Observer& o, double alpha, double beta) {
    return add_rn(mul_rn(add_rn(mul_rn(alpha, alpha), -o.a2), o.cth2), mul_rn(beta, beta));
}

KG_HD void angular_consts(const Observer& o, double lam, double eta, double* up_, double* um_, double* ratio_,
                          double* a2um_) {
    // kgeo/kerr_raytracing_utils.py:276-328, operation for operation
    double up, um;
    const double lam2 = mul_rn(lam, lam);
    if (o.lowspin) {
        up = div_rn(eta, add_rn(eta, lam2));
        um = up;
    } else {
        double dth = mul_rn(0.5, add_rn(1.0, -div_rn(add_rn(eta, lam2), o.a2)));
        double rt = sqrt_rn(add_rn(mul_rn(dth, dth), div_rn(eta, o.a2)));
        up = add_rn(dth, rt);
        um = add_rn(dth, -rt);
    }
    if (o.cth2 - up > 0.0) up = o.cth2;
    if (mul_rn(up, up) > 1.0) up = 1.0;
    if (um == 0.0) um = (eta < 0.0) ? kEP : -kEP;
    if (eta < 0.0 && (o.cth2 - um) < 0.0) um = o.cth2;
    *up_ = up;
    *um_ = um;
    if (o.lowspin) {
        *ratio_ = 0.0;
        *a2um_ = -add_rn(eta, lam2);
    } else {
        *ratio_ = div_rn(up, um);
        *a2um_ = mul_rn(o.a2, um);
    }
}

// Newton polish of a root of r^4 + A r^2 + B r + C (real root, or the upper member x + i y of a complex pair)
KG_HD void polish_real(double A, double B, double C, double* r_) {
    double r = *r_;
    for (int it = 0; it < 2; ++it) {
        double f = fma(fma(fma(r, r, A), r, B), r, C);
        double fp = fma(fma(4.0 * r, r, 2.0 * A), r, B);
        double step = f / fp;
        if (!(fabs(step) <= 1.0e-6 * fmax(1.0, fabs(r)))) break;
        r -= step;
    }
    *r_ = r;
}
KG_HD void polish_complex(double A, double B, double C, double* x_, double* y_) {
    double x = *x_, y = *y_;
    for (int it = 0; it < 2; ++it) {
        // w^2, then Horner in complex arithmetic: f = ((w^2 + A) w + B) w + C, f' = (4 w^2 + 2A) w + B
        double w2r = fma(x, x, -y * y), w2i = 2.0 * x * y;
        double t1r = w2r + A, t1i = w2i;
        double t2r = fma(t1r, x, -t1i * y) + B, t2i = fma(t1r, y, t1i * x);
        double fr = fma(t2r, x, -t2i * y) + C, fi = fma(t2r, y, t2i * x);
        double g1r = fma(4.0, w2r, 2.0 * A), g1i = 4.0 * w2i;
        double gr = fma(g1r, x, -g1i * y) + B, gi = fma(g1r, y, g1i * x);
        double den = fma(gr, gr, gi * gi);
        double sr = fma(fr, gr, fi * gi) / den, si = fma(fi, gr, -fr * gi) / den;
        if (!(fabs(sr) + fabs(si) <= 1.0e-6 * fmax(1.0, fabs(x) + fabs(y)))) break;
        x -= sr;
        y -= si;
    }
    *x_ = x;
    *y_ = y;
}

KG_HD Roots quartic_roots(const Observer& o, double lam, double eta, bool polish = false) {
    // kgeo/kerr_raytracing_utils.py:330-380 in real arithmetic.  The Ferrari resolvent loses digits by cancellation
    // (xi0 = omega_+ + omega_- - A/3), so the reference's own roots carry ~1e-11 relative noise at high spin; to land on
    // the SAME roots the coefficient chain is evaluated with individually rounded operations in NumPy's order.
    Roots R;
    const double a = o.a, a2 = o.a2;
    const double lam2 = mul_rn(lam, lam);
    const double A = add_rn(add_rn(a2, -eta), -lam2);
    const double lma = add_rn(lam, -a);
    const double B = mul_rn(2.0, add_rn(eta, mul_rn(lma, lma)));
    const double C = mul_rn(-a2, eta);
    const double AA = mul_rn(A, A);
    const double P = add_rn(div_rn(-AA, 12.0), -C);
    const double Q = add_rn(mul_rn(div_rn(-A, 3.0), add_rn(div_rn(AA, 36.0), -C)), -div_rn(mul_rn(B, B), 8.0));
    const double disc = add_rn(mul_rn(4.0, mul_rn(mul_rn(P, P), P)), mul_rn(27.0, mul_rn(Q, Q)));
    const double mhq = mul_rn(-0.5, Q);
    const double A3 = div_rn(A, 3.0);
    double xi0;
    // NumPy divides a complex array by a real as  x * (1/y)  (Smith's algorithm with a zero imaginary part), so the
    // reference's disc/108, theta/3 and 0.25 B/z are each a reciprocal followed by a product: two roundings, kept here.
    const double inv108 = div_rn(1.0, 108.0), inv3 = div_rn(1.0, 3.0);
    if (disc >= 0.0) {
        double h = sqrt_rn(mul_rn(disc, inv108));
        xi0 = add_rn(add_rn(cbrt(add_rn(mhq, h)), cbrt(add_rn(mhq, -h))), -A3);
    } else {
        // omega_+/-^3 = -Q/2 +/- i h: the cube roots with the largest real part are complex conjugates
        double h = sqrt_rn(mul_rn(-disc, inv108));
        double mag = hypot(mhq, h);
        double ang = atan2(h, mhq);
        double re = mul_rn(cbrt(mag), cos(mul_rn(ang, inv3)));
        xi0 = add_rn(add_rn(re, re), -A3);
    }
    const double z = sqrt_rn(mul_rn(0.5, xi0));
    const double base = add_rn(mul_rn(-0.5, A), -mul_rn(z, z));
    const double bz = mul_rn(mul_rn(0.25, B), div_rn(1.0, z));
    const double d12 = add_rn(base, bz);
    const double d34 = add_rn(base, -bz);
    R.c12 = d12 < 0.0;
    R.c34 = d34 < 0.0;
    if (R.c12) {
        double s = sqrt(-d12);
        R.re[0] = -z; R.re[1] = -z; R.im[0] = -s; R.im[1] = s;
    } else {
        double s = sqrt(d12);
        R.re[0] = -z - s; R.re[1] = -z + s; R.im[0] = 0.0; R.im[1] = 0.0;
    }
    if (R.c34) {
        double s = sqrt(-d34);
        R.re[2] = z; R.re[3] = z; R.im[2] = -s; R.im[3] = s;
    } else {
        double s = sqrt(d34);
        R.re[2] = z - s; R.re[3] = z + s; R.im[2] = 0.0; R.im[3] = 0.0;
    }
    int region = 0;
    if (R.c12) region = 4;
    else if (R.c34) region = 3;
    else if (R.re[3] < o.rplus) region = 2;
    else if (R.re[3] > o.rplus) region = 1;
    R.region = region;
    // Optional (KGEO_MODE_POLISH_ROOTS): with the class fixed exactly as the reference fixes it, polish the root VALUES by
    // two Newton steps on r^4 + A r^2 + B r + C.  Ferrari's formulas lose up to 6 digits (measured against 50-digit
    // roots: the reference's r2 is off by 4e-10 relative at a=0.94, alpha=3.29, beta=1.20, and tau_tot by 4e-11).
    // Polished roots are accurate to ~1e-16 but no longer share the reference's rounding noise, so PARITY with the
    // reference is better without it (default off).  Pairs closer than 1e-4 are left untouched.
    if (polish) {
        const double scale = fmax(1.0, fabs(z));
        const double s12 = sqrt(fabs(d12)), s34 = sqrt(fabs(d34));
        if (s12 > 1.0e-4 * scale) {
            if (R.c12) polish_complex(A, B, C, &R.re[1], &R.im[1]);
            else { polish_real(A, B, C, &R.re[0]); polish_real(A, B, C, &R.re[1]); }
            if (R.c12) { R.re[0] = R.re[1]; R.im[0] = -R.im[1]; }
        }
        if (s34 > 1.0e-4 * scale) {
            if (R.c34) polish_complex(A, B, C, &R.re[3], &R.im[3]);
            else { polish_real(A, B, C, &R.re[2]); polish_real(A, B, C, &R.re[3]); }
            if (R.c34) { R.re[2] = R.re[3]; R.im[2] = -R.im[3]; }
        }
    }
    return R;
}

// ---- Legendre integrals of an amplitude given by (sin, cos), |phi| <= pi ----------------------------------------
// F, E and NP Pi's; amplitudes beyond pi/2 are reflected (phi = +/-pi + phi_r).  With Kc/RDc/PicTail supplied by
// the caller when known (pass NaN to have them computed on demand).
template <int NP>
KG_NOINLINE void legendre_sc(double s, double c, double m, const double* n, double* F, double* E, double* Pi) {
    double j = 0.0;
    if (c < 0.0) {
        j = (s >= 0.0) ? 1.0 : -1.0;
        s = -s;
        c = -c;
    }
    double c2 = c * c, y = fma(-m * s, s, 1.0);
    PiPrep pp[NP > 0 ? NP : 1];
    double pa[NP > 0 ? NP : 1], rj[NP > 0 ? NP : 1];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        pp[i] = pi_prepare(n[i], s, c2, y);
        pa[i] = pp[i].p;
    }
    double rf, rd;
    carlson_bundle<NP, true, true>(c2, y, pa, &rf, &rd, rj);
    double Fr = s * rf;
    double Er = Fr - (m * (1.0 / 3.0)) * cube(s) * rd;
    double Kc = 0.0, RDc = 0.0;
    if (j != 0.0) {
        Kc = ellip_K(m);
        RDc = ellip_RDc(m);
    }
    *F = (j != 0.0) ? fma(2.0 * j, Kc, Fr) : Fr;
    *E = (j != 0.0) ? fma(2.0 * j, Kc - (m * (1.0 / 3.0)) * RDc, Er) : Er;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double v = Fr + pi_finish(n[i], s, c2, y, Fr, pp[i], rj[i]);
        if (j != 0.0) v += 2.0 * j * (Kc + ellip_PicTail(n[i], m));
        Pi[i] = v;
    }
}

KG_HD double legendre_F_sc(double s, double c, double m) {
    double F, E;
    legendre_sc<0>(s, c, m, nullptr, &F, &E, nullptr);
    return F;
}
}  // namespace kg
#include "kg_theta.cuh"
namespace kg {

// ---- total Mino time from given roots (kgeo/kerr_raytracing_utils.py:382-548) ------------------------------------------
// Stand-alone version for the helper API mino_total(a, r_o, eta, r1..r4); ray_setup computes the same quantity in
// passing from its own case constants.
KG_HD double tau_total_of(const Observer& o, const Roots& R) {
    const double rp = o.rplus, r_o = o.r_o;
    if (R.c12) {
        double a1 = R.im[3], b1 = R.re[3], a2 = R.im[1], b2 = R.re[1];
        double CC = sqrt(sq(a1 - a2) + sq(b1 - b2)), DD = sqrt(sq(a1 + a2) + sq(b1 - b2));
        double k4 = (4.0 * CC * DD) / sq(CC + DD);
        double g0 = sqrt((4.0 * a2 * a2 - sq(CC - DD)) / (sq(CC + DD) - 4.0 * a2 * a2));
        double x4rp = (rp + b1) / a2;
        double ig = 1.0 / sqrt(1.0 + g0 * g0);
        double so, co;
        if (o.ro_inf) {
            so = ig;
            co = -g0 * ig;
        } else {
            double x4ro = (r_o + b1) / a2, ix = 1.0 / sqrt(1.0 + x4ro * x4ro);
            so = (x4ro + g0) * ix * ig;
            co = (1.0 - x4ro * g0) * ix * ig;
        }
        double ixp = 1.0 / sqrt(1.0 + x4rp * x4rp);
        double sh = (x4rp + g0) * ixp * ig, ch = (1.0 - x4rp * g0) * ixp * ig;
        double pref = 2.0 / (CC + DD);
        return pref * legendre_F_sc(so, co, k4) - pref * legendre_F_sc(sh, ch, k4);
    }
    if (R.c34) {
        double a1 = R.im[3], b1 = R.re[3], rr1 = R.re[0], rr2 = R.re[1];
        double rr21 = rr2 - rr1, rrp1 = rp - rr1, rrp2 = rp - rr2;
        double AA = sqrt(a1 * a1 + sq(b1 - rr2)), BB = sqrt(a1 * a1 + sq(b1 - rr1));
        double k3 = (sq(AA + BB) - rr21 * rr21) / (4.0 * AA * BB);
        double x3rp = (AA * rrp1 - BB * rrp2) / (AA * rrp1 + BB * rrp2);
        double x3ro = o.ro_inf ? (AA - BB) / (AA + BB)
                               : (AA * (r_o - rr1) - BB * (r_o - rr2)) / (AA * (r_o - rr1) + BB * (r_o - rr2));
        double pref = 1.0 / sqrt(AA * BB);
        double so = sqrt(fmax(0.0, (1.0 - x3ro) * (1.0 + x3ro))), sh = sqrt(fmax(0.0, (1.0 - x3rp) * (1.0 + x3rp)));
        return pref * legendre_F_sc(so, x3ro, k3) - pref * legendre_F_sc(sh, x3rp, k3);
    }
    double q1 = R.re[0], q2 = R.re[1], q3 = R.re[2], q4 = R.re[3];
    double rr31 = q3 - q1, rr32 = q3 - q2, rr41 = q4 - q1, rr42 = q4 - q2;
    double k2 = (rr32 * rr41) / (rr31 * rr42);
    double x2ro = o.ro_inf ? sqrt(rr31 / rr41) : sqrt((rr31 * (r_o - q4)) / (rr41 * (r_o - q3)));
    double pref = 2.0 / sqrt(rr31 * rr42);
    double I2ro = pref * legendre_F_sc(x2ro, sqrt(fmax(0.0, (1.0 - x2ro) * (1.0 + x2ro))), k2);
    if (q3 < rp) {
        double x2rp = sqrt((rr31 * (rp - q4)) / (rr41 * (rp - q3)));
        return I2ro - pref * legendre_F_sc(x2rp, sqrt(fmax(0.0, (1.0 - x2rp) * (1.0 + x2rp))), k2);
    }
    if (q3 > rp) return 2.0 * I2ro;
    return 0.0;
}

// ---- GL19a auxiliary integrals ------------------------------------------------------------------------------------
// R1, R2 (B62, B65; kgeo/kerr_raytracing_ana.py:666-683).  sn, cn, dn = sin, cos, sqrt(1-j sin^2) of the amplitude.
// FAST = true (per-sample code): divisions by the 1-ulp fdiv of kg_special.cuh; FAST = false (per-ray setup): IEEE.
template <bool FAST>
KG_HD double qdiv(double a, double b) {
    return FAST ? fdiv(a, b) : a / b;
}
template <bool FAST = false>
KG_HD double R1_of(double al, double p1, double nn, double sn, double dn, double Pi) {
    double f1 = 0.5 * p1 * kg_log(fabs(qdiv<FAST>(p1 * dn + sn, p1 * dn - sn)));
    return qdiv<FAST>(Pi - al * f1, 1.0 - al * al);
}
template <bool FAST = false>
KG_HD double R2_of(double al, double j, double nn, double sn, double cn, double dn, double F, double E, double R1) {
    double al2 = al * al, iden = qdiv<FAST>(1.0, j + (1.0 - j) * al2);
    double R2 = qdiv<FAST>(F - (al2 * iden) * (E - qdiv<FAST>(al * sn * dn, 1.0 + al * cn)), al2 - 1.0);
    return R2 + (2.0 * j - nn) * R1 * iden;
}
// S1, S2 (B92, B93; kgeo/kerr_raytracing_ana.py:638-663)
template <bool FAST = false>
KG_HD double S1_of(double al, double p2, double dn, double F, double Pi) {
    double al2 = al * al;
    double f2 = 0.5 * p2 * kg_log(fabs(qdiv<FAST>((1.0 - p2) * (1.0 + p2 * dn), (1.0 + p2) * (1.0 - p2 * dn))));
    return qdiv<FAST>(F + al2 * Pi - al * f2, 1.0 + al2);
}
template <bool FAST = false>
KG_HD double S2_of(double al, double j, double sn, double cn, double dn, double F, double E, double S1) {
    double al2 = al * al, al2p1 = 1.0 + al2, al2j = 1.0 - j + al2;
    // (al - tan)/(1 + al tan) written with sn, cn so that cn -> 0 is harmless
    double S2a = (1.0 - j) * F + al2 * E + qdiv<FAST>(al2 * dn * (al * cn - sn), cn + al * sn) - al2 * al;
    double i1 = qdiv<FAST>(1.0, al2p1), i2 = qdiv<FAST>(1.0, al2j);
    S2a = -S2a * (i1 * i2);
    return S2a + S1 * (i1 + (1.0 - j) * i2);
}

// ---- theta blocks of a motion (setup) -------------------------------------------------------------------------------------
// Builds base + NP Pi blocks at w[slot...]; returns true when the theta path is usable for every integral of the motion.
template <class P>
KG_NOINLINE bool theta_make_motion(int NP, const ThetaRay& r, const double* n, const double* nmm, const double* omn, double pmax, double edge, P blk) {
    const double m = r.m;
    int J = theta_make_base(r, blk);
    bool ok = r.ok != 0;
    for (int i = 0; i < NP; ++i) {
        int Ji = theta_make_pi(r, n[i], nmm ? nmm[i] : n[i] - m, omn ? omn[i] : 1.0 - n[i], pmax, edge, blk + TB_SIZE + i * TP_SIZE);
        if (Ji < 0) ok = false;
        else if (Ji > J) J = Ji;
    }
    if (J > kThJM) J = kThJM;
    blk[TB_J] = (double)J;
    blk[TB_OK] = ok ? 1.0 : 0.0;
    return ok;
}

// polar motion: Legendre parameter k < 0 of the reference, Jacobi parameter m = k/(k-1); with sin(phi) = sqrt(1-m) sd(w|m):
//   Pi(n; phi | k) = sqrt(1-m) [ (m/n') w + (1 - m/n') Pi(n'; am w | m) ],  n' = m + n (1 - m)   (m < n' < 1: circular case)
//   (F - E)(phi | k) / k = sqrt(1-m) [ (1 + (E/K - 1)/m) w + Z(w)/m - sn w cn w / dn w ]
template <class WP>
KG_NOINLINE void polar_theta_setup(const Observer& o, double m, double sq1m, double n, double D3, WP w, int* flags) {
    const double Km = w[T_K];                      // K(m) of the first setup kernel
    const ThetaRay tr = theta_ray(m, Km);
    const double npr = fma(n, 1.0 - m, m), nmm = n * (1.0 - m), omn = (1.0 - n) * (1.0 - m);
    bool ok = theta_make_motion(1, tr, &npr, &nmm, &omn, kThetaPMaxPolar, kThetaEdgePolar, w + T_TH);
    const double A = w[T_TH + TB_SIZE + TP_A], P = w[T_TH + TB_SIZE + TP_P];
    const double mn = m / npr;
    w[T_TH_SLOPE] = D3 * sq1m * fma(1.0 - mn, A, mn);
    w[T_TH_COEF] = D3 * sq1m * (1.0 - mn) * P;
    w[T_TH_G] = tr.ok ? 1.0 - tr.RDc / (3.0 * Km) : 0.0;      // only read on the theta path
    w[T_TH_IM] = 1.0 / m;
    if (ok && !o.lowspin && m > 1.0e-6 && w[T_TH + TB_SIZE + TP_KIND] == 2.0) *flags |= F_TH_THETA;
}

template <class WP>
KG_HD void radial_theta_setup(double k, const double* n, WP w, int* flags) {
    const ThetaRay tr = theta_ray(k, w[R_KC], w[R_RDC]);      // K(k), RD(0, 1-k, 1) of the first setup kernel
    if (!theta_make_motion(3, tr, n, nullptr, nullptr, kThetaPMax, kThetaEdge, w + R_TH)) return;
    *flags |= F_R_THETA;
    const double k0 = w[R_TH + TB_SIZE + TP_KIND], k1 = w[R_TH + TB_SIZE + TP_SIZE + TP_KIND],
                 k2 = w[R_TH + TB_SIZE + 2 * TP_SIZE + TP_KIND];
    if (k0 == 1.0 && k1 == 0.0 && k2 == 0.0) *flags |= F_R_K100;
    if (k0 == 1.0 && k1 == 1.0 && k2 == 1.0) *flags |= F_R_K111;
}

// ---- ray setup -------------------------------------------------------------------------------------------------------
// any_eta_pos / any_eta_nonpos reproduce the reference's whole-array guards `np.any(eta>0)` / `np.any(eta<=0)`
// (kgeo/kerr_raytracing_ana.py:217,300): they are properties of the CALL, computed once by the host wrapper.
enum SetupMode { M_ANY_ETA_POS = 1, M_ANY_ETA_NONPOS = 2, M_CLAMP_LOWSPIN_ETA = 4, M_INPUT_LAM_ETA = 8, M_POLISH_ROOTS = 16, M_CARLSON_ONLY = 32 };
// DEFER = true leaves the theta-function blocks to theta_setup (second kernel of the two-kernel setup).
// PHI_T = false skips every constant that only the phi / t / sigma integrals need (image and lensing paths).
// LADDER = false skips the stored AGM ladders (their slots are indexed dynamically, which would push a register-resident
// workspace into local memory): the image / lensing kernels evaluate only a few Jacobi functions per ray and run the
// AGM descent inline instead.
// ray_setup_conserved takes the conserved quantities themselves (lam, eta, and the sign s_o of the observer's polar momentum =
// sign(beta)): what th_integrate / r_integrate receive from their callers (kgeo/kerr_raytracing_ana.py:195,368); ray_setup derives
// them from the image coordinates (:76-79).
// PART splits the work for the track setup's two launches (each half of the ~15 k instructions stays in the instruction caches;
// the halves share nothing but lam and eta): 0 = everything, 1 = conserved quantities + polar motion, 2 = radial motion (the
// caller passes the flags part 1 left in *flags_out; every slot and flag is written by exactly one part).
KG_HD constexpr bool slot_is_radial(int i) {
    return (i >= R_K && i <= R_Q10) || i == R_I2K || (i >= R_C0 && i <= R_RDC) || (i >= R_TH && i <= R_TH_END) || (i >= R_LAD && i <= R_LAD_END) ||
           i == S_TAUTOT;
}
template <bool PHI_T, bool LADDER = true, bool DEFER = false, class P = double*, int PART = 0>
KG_HD void ray_setup_conserved(const Observer& o, double lam, double eta, bool s_pos, int mode, P w, int* flags_out) {
    static_assert(PART == 0 || LADDER, "the split setup is for the grouped track workspace");
    const int any_eta_pos = mode & M_ANY_ETA_POS, any_eta_nonpos = mode & M_ANY_ETA_NONPOS;
    // Two-kernel setup (DEFER): the AGM ladders are only read by the Carlson fallback, so they are built where the theta path
    // is turned down (theta_setup, or at the end of this function for motions that never were candidates) instead of for every
    // ray, and neither they nor the theta blocks are zero-filled first: 190 of the 257 slots, which the first version
    // wrote twice (ncu: 3.1 GB of DRAM writes for a 2.15 GB workspace).  Slots a ray does not use stay undefined.
    const bool defer_ladders = DEFER && !(mode & M_CARLSON_ONLY);
    if constexpr (!LADDER) {
        // register-resident workspace of the image / lensing kernels: static indices only (a rolled loop would index the array
        // dynamically and put all of it in local memory)
#pragma unroll
        for (int i = 0; i < NW_LITE; ++i) w[i] = 0.0;
    } else {
        for (int i = 0; i < NW; ++i) {
            if (DEFER && ((i >= T_LAD && i <= R_LAD_END) || (i >= T_TH && i <= R_TH_END))) continue;
            if (PART != 0 && slot_is_radial(i) != (PART == 2)) continue;
            w[i] = 0.0;
        }
    }
    int flags = (PART == 2) ? *flags_out : 0;
    if (o.lowspin && (mode & M_CLAMP_LOWSPIN_ETA) && eta < 0.0) eta = kEP;   // kgeo/kerr_raytracing_ana.py:80-82 (tracks only)
  if constexpr (PART != 2) {
    if (o.lowspin) flags |= F_LOWSPIN;
    w[S_LAM] = lam;
    w[S_ETA] = eta;
    double s_o = s_pos ? 1.0 : -1.0;
    if (s_pos) flags |= F_SPOS;
    if (o.th_o <= kHalfPi) flags |= F_HPOS;

    double up, um, ratio, a2um;
    angular_consts(o, lam, eta, &up, &um, &ratio, &a2um);

    // ---------------- polar motion (kgeo/kerr_raytracing_ana.py:216-363) ----------------
    if (any_eta_pos && eta >= 0.0) {
        flags |= F_TH_ORD;
        double k = ratio;                        // <= 0
        double pref = 1.0 / sqrt(-a2um);
        double so = o.cth / sqrt(up);            // sin of the observer amplitude, in [-1,1]
        double co = sqrt(fmax(0.0, (1.0 - so) * (1.0 + so)));
        double F0, E0 = 0.0, Pi0 = 0.0;
        if (PHI_T) legendre_sc<1>(so, co, k, &up, &F0, &E0, &Pi0);
        else F0 = legendre_F_sc(so, co, k);
        double m = k / (k - 1.0);
        double sq1m = sqrt(1.0 - m);
        double Km, RDc = 0.0, PicT = 0.0;
        ellip_complete<0, true, false>(m, nullptr, &Km, nullptr, nullptr);
        double Kk = sq1m * Km;                   // K(k) for the negative parameter
        if (PHI_T) ellip_complete<1, false, true>(k, &up, nullptr, &RDc, &PicT);   // RD(0, 1-k, 1); Pi(u+|k) - K(k), +inf when u+ == 1
        w[T_W] = sqrt(-a2um);
        w[T_X0] = w[T_W] * s_o * (-pref * F0);
        w[T_M] = m;
        w[T_SQ] = sq1m;
        w[T_ISQ] = 1.0 / sq1m;
        w[T_I2K] = (Km < kInf) ? 0.5 / Km : 0.0;
        if (LADDER && !defer_ladders) jacobi_make_ladder(m, w + T_LAD);
        w[T_K] = Km;
        w[T_KPAR] = k;
        w[T_A] = -s_o * sqrt(up);
        w[T_N] = up;
        w[T_D2] = 2.0 * pref * PicT;
        w[T_D3] = pref;
        w[T_GPH0] = s_o * (-pref * Pi0);
        double ct = up * pref * (1.0 / 3.0);
        w[T_C1] = ct;
        w[T_C2] = 2.0 * ct * RDc;
        w[T_C3] = 0.0;
        // Gt_o = 2 u+ pref (E - F)/(2k) = -(u+ pref/3) s^3 RD, with E - F = -(k/3) s^3 RD
        if (PHI_T) {
            double EmF_over_k = (k != 0.0) ? (E0 - F0) / k : -(1.0 / 3.0) * cube(so) * carlson_rd(co * co, 1.0, 1.0);
            w[T_GT0] = s_o * (up * pref * EmF_over_k);
        }
        w[T_F0] = F0;
        w[T_KC] = Kk;
        if (PHI_T && LADDER) {
            if (DEFER) { w[T_TH_SLOPE] = up; w[T_TH_COEF] = pref; flags |= F_TH_PENDING; }
            else polar_theta_setup(o, m, sq1m, up, pref, w, &flags);
        }
    } else if (any_eta_nonpos && eta < 0.0) {
        flags |= F_TH_VORT;
        double h = (o.th_o <= kHalfPi) ? 1.0 : -1.0;
        double prefA = 1.0 / sqrt(a2um);
        double prefB = prefA / (1.0 - um);
        double prefC = sqrt(um / o.a2);
        double so = sqrt((o.cth2 - um) / (up - um));
        double co = sqrt(fmax(0.0, (1.0 - so) * (1.0 + so)));
        double k = 1.0 - ratio;
        double nch = (up - um) / (1.0 - um);
        double F0, E0 = 0.0, Pi0 = 0.0;
        if (PHI_T) legendre_sc<1>(so, co, k, &nch, &F0, &E0, &Pi0);
        else F0 = legendre_F_sc(so, co, k);
        double m = k / (k - 1.0);
        double sq1m = sqrt(1.0 - m);
        double Km, Ek = 0.0, PicT = 0.0;
        ellip_complete<0, true, false>(m, nullptr, &Km, nullptr, nullptr);
        double Kk = sq1m * Km;
        if (PHI_T) {
            double Kk_, RDk;
            ellip_complete<1, true, true>(k, &nch, &Kk_, &RDk, &PicT);
            Ek = Kk_ - (k * (1.0 / 3.0)) * RDk;  // E(k)
        }
        w[T_W] = sqrt(a2um);
        w[T_X0] = w[T_W] * s_o * (-h * prefA * F0);
        w[T_M] = m;
        w[T_SQ] = sq1m;
        w[T_ISQ] = 1.0 / sq1m;
        w[T_I2K] = (Km < kInf) ? 0.5 / Km : 0.0;
        if (LADDER && !defer_ladders) jacobi_make_ladder(m, w + T_LAD);
        w[T_K] = Km;
        w[T_KPAR] = k;
        w[T_A] = um;
        w[T_B] = up - um;
        w[T_N] = nch;
        w[T_D2] = 2.0 * prefB * PicT;
        w[T_D3] = prefB;
        w[T_GPH0] = s_o * (-h * prefB * Pi0);
        w[T_C1] = -prefC * k * (1.0 / 3.0);
        w[T_C2] = 2.0 * prefC * (Ek - Kk);
        w[T_C3] = prefC;
        w[T_GT0] = s_o * (-h * prefC * E0);
        if (PHI_T && LADDER) {
            if (DEFER) { w[T_TH_SLOPE] = nch; w[T_TH_COEF] = prefB; flags |= F_TH_PENDING; }
            else polar_theta_setup(o, m, sq1m, nch, prefB, w, &flags);
        }
    }

  }
  if constexpr (PART != 1) {
    // ---------------- radial motion (kgeo/kerr_raytracing_ana.py:368-633, utils :382-548) ----------------
    Roots R = quartic_roots(o, lam, eta, (mode & M_POLISH_ROOTS) != 0);
    flags |= (R.region << F_REGION_SHIFT);
    if (!R.c34 && R.re[3] > o.rplus && !(eta < 0.0)) flags |= F_OUTSIDE;   // utils :817-824
    const double rp = o.rplus, rm = o.rminus, r_o = o.r_o;
    double tau_tot = 0.0;
    double th_k = -1.0, th_n[3] = {0.0, 0.0, 0.0};   // Jacobi parameter and Pi characteristics for the theta-function blocks

    if (R.c12) {  // ---- case 4 ----
        flags |= F_CASE4;
        double a1 = R.im[3], b1 = R.re[3], a2 = R.im[1], b2 = R.re[1];
        double CC = sqrt(sq(a1 - a2) + sq(b1 - b2));
        double DD = sqrt(sq(a1 + a2) + sq(b1 - b2));
        double k4 = (4.0 * CC * DD) / sq(CC + DD);
        double x4rp = (rp + b1) / a2, x4rm = (rm + b1) / a2;
        double g0 = sqrt((4.0 * a2 * a2 - sq(CC - DD)) / (sq(CC + DD) - 4.0 * a2 * a2));
        double gp = (g0 * x4rp - 1.0) / (g0 + x4rp);
        double gm = (g0 * x4rm - 1.0) / (g0 + x4rm);
        double pref = 2.0 / (CC + DD);
        // aux = atan(x4ro) + atan(g0) as (sin, cos)
        double ig = 1.0 / sqrt(1.0 + g0 * g0);
        double sa, ca;
        if (o.ro_inf) {
            sa = ig;
            ca = -g0 * ig;
        } else {
            double x4ro = (r_o + b1) / a2;
            double ix = 1.0 / sqrt(1.0 + x4ro * x4ro);
            sa = (x4ro + g0) * ix * ig;
            ca = (1.0 - x4ro * g0) * ix * ig;
        }
        double nn[3] = {1.0 + g0 * g0, 1.0 + gp * gp, 1.0 + gm * gm};
        double al[3] = {g0, gp, gm};
        double Fo, Eo = 0.0, Pio[3] = {0.0, 0.0, 0.0};
        if (PHI_T) legendre_sc<3>(sa, ca, k4, nn, &Fo, &Eo, Pio);
        else Fo = legendre_F_sc(sa, ca, k4);
        w[R_K] = k4;
        if (LADDER && !defer_ladders) jacobi_make_ladder(k4, w + R_LAD);
        th_k = k4; th_n[0] = nn[0]; th_n[1] = nn[1]; th_n[2] = nn[2];
        w[R_SC] = 1.0 / pref;
        w[R_X0] = Fo;                          // X4 = (1/pref)(-tau + pref*Fo) = Fo - tau/pref
        if (PHI_T) {
            double dno = sqrt(fma(-k4 * sa, sa, 1.0));
            double Kc_, RDc_, PicT_[3];
            ellip_complete<3>(k4, nn, &Kc_, &RDc_, PicT_);      // K, RD(0, 1-k, 1) and the three Pi(n_i|k) - K(k): one track
            w[R_KC] = Kc_;
            w[R_I2K] = (Kc_ < kInf) ? 0.5 / Kc_ : 0.0;
            w[R_RDC] = RDc_;
            w[R_EC] = Kc_ - (k4 * (1.0 / 3.0)) * RDc_;            // E(k)
            for (int i = 0; i < 3; ++i) {
                double p2 = sqrt((1.0 + al[i] * al[i]) / (1.0 - k4 + al[i] * al[i]));
                w[R_N0 + i] = nn[i];
                w[R_PC0 + i] = PicT_[i];
                w[R_Q5 + i] = p2;
                w[R_PO0 + i] = S1_of(al[i], p2, dno, Fo, Pio[i]);
            }
            w[R_EO] = S2_of(g0, k4, sa, ca, dno, Fo, Eo, w[R_PO0]);
            // per-ray factors of the per-sample algebra (radial_eval, case 4)
            for (int i = 0; i < 3; ++i) {
                double p2 = w[R_Q5 + i];
                w[R_C0 + i] = 1.0 / (1.0 + al[i] * al[i]);
                w[R_C3 + i] = 0.5 * log(fabs((1.0 - p2) / (1.0 + p2)));
            }
            double a2g = a2 / g0, opg = 1.0 + g0 * g0;
            w[R_C6] = a2g;
            w[R_C7] = (pref * opg) / (g0 * g0 + g0 * x4rp);
            w[R_C8] = (pref * opg) / (g0 * g0 + g0 * x4rm);
            w[R_C9] = g0 / (a2 - a2 * g0 * x4rp);
            w[R_C10] = g0 / (a2 - a2 * g0 * x4rm);
            w[R_C11] = 1.0 / (1.0 - k4 + g0 * g0);
        }
        w[R_Q0] = a2; w[R_Q1] = b1; w[R_Q2] = g0; w[R_Q3] = gp; w[R_Q4] = gm;
        w[R_Q8] = pref; w[R_Q9] = x4rp; w[R_Q10] = x4rm;
        // tau_tot (utils :412-442): horizon amplitude atan(x4rp) + atan(g0)
        double ixp = 1.0 / sqrt(1.0 + x4rp * x4rp);
        double sh = (x4rp + g0) * ixp * ig, ch = (1.0 - x4rp * g0) * ixp * ig;
        tau_tot = pref * Fo - pref * legendre_F_sc(sh, ch, k4);
    }
    if (R.c34 && !R.c12) {  // ---- case 3 ----
        flags |= F_CASE3;
        double a1 = R.im[3], b1 = R.re[3], rr1 = R.re[0], rr2 = R.re[1];
        double rr21 = rr2 - rr1, rrp1 = rp - rr1, rrp2 = rp - rr2, rrm1 = rm - rr1, rrm2 = rm - rr2;
        double AA = sqrt(a1 * a1 + sq(b1 - rr2));
        double BB = sqrt(a1 * a1 + sq(b1 - rr1));
        double k3 = (sq(AA + BB) - rr21 * rr21) / (4.0 * AA * BB);
        double x3rp = (AA * rrp1 - BB * rrp2) / (AA * rrp1 + BB * rrp2);
        double x3rm = (AA * rrm1 - BB * rrm2) / (AA * rrm1 + BB * rrm2);
        double x3ro = o.ro_inf ? (AA - BB) / (AA + BB)
                               : (AA * (r_o - rr1) - BB * (r_o - rr2)) / (AA * (r_o - rr1) + BB * (r_o - rr2));
        double al[3] = {(BB + AA) / (BB - AA), -1.0 / x3rp, -1.0 / x3rm};
        double nn[3], p1[3];
        for (int i = 0; i < 3; ++i) {
            double al2 = al[i] * al[i];
            nn[i] = al2 / (al2 - 1.0);
            p1[i] = sqrt((al2 - 1.0) / (k3 + (1.0 - k3) * al2));
        }
        double sqAB = sqrt(AA * BB);
        double ca = x3ro, sa = sqrt(fmax(0.0, (1.0 - ca) * (1.0 + ca)));
        double Fo, Eo = 0.0, Pio[3] = {0.0, 0.0, 0.0};
        if (PHI_T) legendre_sc<3>(sa, ca, k3, nn, &Fo, &Eo, Pio);
        else Fo = legendre_F_sc(sa, ca, k3);
        w[R_K] = k3;
        if (LADDER && !defer_ladders) jacobi_make_ladder(k3, w + R_LAD);
        th_k = k3; th_n[0] = nn[0]; th_n[1] = nn[1]; th_n[2] = nn[2];
        w[R_SC] = sqAB;
        w[R_X0] = Fo;                          // X3 = sqrt(AB)(-tau + Fo/sqrt(AB))
        if (PHI_T) {
            double dno = sqrt(fma(-k3 * sa, sa, 1.0));
            double Kc_, RDc_, PicT_[3];
            ellip_complete<3>(k3, nn, &Kc_, &RDc_, PicT_);      // K, RD(0, 1-k, 1) and the three Pi(n_i|k) - K(k): one track
            w[R_KC] = Kc_;
            w[R_I2K] = (Kc_ < kInf) ? 0.5 / Kc_ : 0.0;
            w[R_RDC] = RDc_;
            w[R_EC] = Kc_ - (k3 * (1.0 / 3.0)) * RDc_;            // E(k)
            for (int i = 0; i < 3; ++i) {
                w[R_N0 + i] = nn[i];
                w[R_PC0 + i] = PicT_[i];
                w[R_Q4 + i] = al[i];
                w[R_Q7 + i] = p1[i];
                w[R_PO0 + i] = R1_of(al[i], p1[i], nn[i], sa, dno, Pio[i]);
            }
            w[R_EO] = R2_of(al[0], k3, nn[0], sa, ca, dno, Fo, Eo, w[R_PO0]);
            // per-ray factors of the per-sample algebra (radial_eval, case 3)
            for (int i = 0; i < 3; ++i) w[R_C0 + i] = 1.0 / (1.0 - al[i] * al[i]);
            double fac = 2.0 * rr21 * sqAB;
            w[R_C3] = fac / (BB * BB - AA * AA);
            w[R_C4] = fac / (BB * rrp2 - AA * rrp1);
            w[R_C5] = fac / (BB * rrm2 - AA * rrm1);
            w[R_C6] = (BB * rr2 + AA * rr1) / (BB + AA);
            w[R_C7] = -1.0 / (BB * rrp2 + AA * rrp1);
            w[R_C8] = -1.0 / (BB * rrm2 + AA * rrm1);
            w[R_C9] = 1.0 / (k3 + (1.0 - k3) * al[0] * al[0]);
            w[R_C10] = 1.0 / (al[0] * al[0] - 1.0);
        }
        w[R_Q0] = rr1; w[R_Q1] = rr2; w[R_Q2] = AA; w[R_Q3] = BB; w[R_Q10] = sqAB;
        // tau_tot (utils :445-482): horizon amplitude acos(x3rp)
        double sh = sqrt(fmax(0.0, (1.0 - x3rp) * (1.0 + x3rp)));
        tau_tot = Fo / sqAB - legendre_F_sc(sh, x3rp, k3) / sqAB;
    }
    if (!R.c34 && !R.c12) {  // ---- case 2 (regions I and II) ----
        flags |= F_CASE2;
        double q1 = R.re[0], q2 = R.re[1], q3 = R.re[2], q4 = R.re[3];
        double rr31 = q3 - q1, rr32 = q3 - q2, rr41 = q4 - q1, rr42 = q4 - q2, rr43 = q4 - q3;
        double rrp3 = rp - q3, rrm3 = rm - q3, rrp4 = rp - q4, rrm4 = rm - q4;
        double k2 = (rr32 * rr41) / (rr31 * rr42);
        double x2ro = o.ro_inf ? sqrt(rr31 / rr41) : sqrt((rr31 * (r_o - q4)) / (rr41 * (r_o - q3)));
        double root = sqrt(rr31 * rr42);
        double nn[3] = {rr41 / rr31, (rrp3 * rr41) / (rrp4 * rr31), (rrm3 * rr41) / (rrm4 * rr31)};
        double sa = x2ro, ca = sqrt(fmax(0.0, (1.0 - sa) * (1.0 + sa)));
        double Fo, Eo = 0.0, Pio[3] = {0.0, 0.0, 0.0};
        if (PHI_T) legendre_sc<3>(sa, ca, k2, nn, &Fo, &Eo, Pio);
        else Fo = legendre_F_sc(sa, ca, k2);
        w[R_K] = k2;
        if (LADDER && !defer_ladders) jacobi_make_ladder(k2, w + R_LAD);
        th_k = k2; th_n[0] = nn[0]; th_n[1] = nn[1]; th_n[2] = nn[2];
        w[R_SC] = 0.5 * root;
        w[R_X0] = Fo;                          // X2 = 0.5 root (-tau + 2 Fo/root)
        if (PHI_T) {
            double Kc_, RDc_, PicT_[3];
            ellip_complete<3>(k2, nn, &Kc_, &RDc_, PicT_);      // K, RD(0, 1-k, 1) and the three Pi(n_i|k) - K(k): one track
            w[R_KC] = Kc_;
            w[R_I2K] = (Kc_ < kInf) ? 0.5 / Kc_ : 0.0;
            w[R_RDC] = RDc_;
            w[R_EC] = Kc_ - (k2 * (1.0 / 3.0)) * RDc_;            // E(k)
            for (int i = 0; i < 3; ++i) {
                w[R_N0 + i] = nn[i];
                w[R_PC0 + i] = PicT_[i];
                w[R_PO0 + i] = Pio[i];
            }
            w[R_EO] = Eo;
        }
        w[R_Q0] = q1; w[R_Q1] = q2; w[R_Q2] = q3; w[R_Q3] = q4; w[R_Q4] = root;
        if (PHI_T) {   // per-ray factors of the case-2 assembly, so that the per-sample code has no division by a constant
            double c = 2.0 / root;
            w[R_Q6] = c;
            w[R_Q7] = c * (rr43 / (rrp3 * rrp4));
            w[R_Q8] = c * (rr43 / (rrm3 * rrm4));
            w[R_Q9] = 1.0 / rrp3;
            w[R_Q10] = 1.0 / rrm3;
        }
        double Rpot_o = (r_o - q1) * (r_o - q2) * (r_o - q3) * (r_o - q4);
        w[R_Q5] = sqrt(Rpot_o) / (r_o - q3);   // H_o (NaN for r_o = inf, as in the reference)
        // tau_tot (utils :485-546): note the reference tests r3 (not r4) against r_+
        double Iro = 2.0 * Fo / root;
        if (q3 < rp) {
            double x2rp = sqrt((rr31 * (rp - q4)) / (rr41 * (rp - q3)));
            double ch = sqrt(fmax(0.0, (1.0 - x2rp) * (1.0 + x2rp)));
            tau_tot = Iro - 2.0 * legendre_F_sc(x2rp, ch, k2) / root;
        } else if (q3 > rp) {
            tau_tot = 2.0 * Iro;
        }
    }
    w[S_TAUTOT] = tau_tot;
    if (PHI_T && LADDER && !(mode & M_CARLSON_ONLY) && th_k > 0.0) {
        if (DEFER) flags |= F_R_PENDING;            // k and the characteristics are in w[R_K], w[R_N0..2]
        else radial_theta_setup(th_k, th_n, w, &flags);
    }
  }
    if (mode & M_CARLSON_ONLY) flags &= ~(F_TH_THETA | F_R_THETA | F_TH_PENDING | F_R_PENDING);
    if (LADDER && defer_ladders) {      // motions theta_setup will not visit keep the Carlson path: their ladders are built here
        if (PART != 2 && !(flags & F_TH_PENDING) && (flags & (F_TH_ORD | F_TH_VORT))) jacobi_make_ladder(w[T_M], w + T_LAD);
        if (PART != 1 && !(flags & F_R_PENDING)) jacobi_make_ladder(w[R_K], w + R_LAD);
    }
    *flags_out = flags;
}

template <bool PHI_T, bool LADDER = true, bool DEFER = false, class P = double*, int PART = 0>
KG_HD void ray_setup(const Observer& o, double alpha, double beta, int mode, P w, int* flags_out) {
    ray_setup_conserved<PHI_T, LADDER, DEFER, P, PART>(o, -alpha * o.sth, conserved_eta(o, alpha, beta), beta >= 0.0, mode, w, flags_out);
}

// Second half of the two-kernel setup: builds the theta blocks ray_setup<.., DEFER = true> left pending.  Splitting the
// setup keeps each kernel's code small enough for the instruction caches (every warp walks the setup code once, front to
// back: with everything in one 16 k-instruction kernel the dominant stall was instruction fetch).
template <class P>
KG_HD void theta_setup(const Observer& o, P w, int* flags_io) {
    int flags = *flags_io;
    if (flags & F_TH_PENDING) {
        flags &= ~F_TH_PENDING;
        polar_theta_setup(o, w[T_M], w[T_SQ], w[T_TH_SLOPE], w[T_TH_COEF], w, &flags);
        if (!(flags & F_TH_THETA)) jacobi_make_ladder(w[T_M], w + T_LAD);      // Carlson fallback: needs the AGM ladder
    }
    if (flags & F_R_PENDING) {
        flags &= ~F_R_PENDING;
        const double n[3] = {w[R_N0], w[R_N1], w[R_N2]};
        radial_theta_setup(w[R_K], n, w, &flags);
        if (!(flags & F_R_THETA)) jacobi_make_ladder(w[R_K], w + R_LAD);
    }
    *flags_io = flags;
}

// view of the workspace starting at a base slot (used for the ladders)
template <class W>
struct SlotView {
    W w;
    int base;
    KG_HD double operator()(int i) const { return w(base + i); }
};

// ---- per-sample evaluation ---------------------------------------------------------------------------------------
struct Point {
    double t, r, th, ph, sig;
};

// reduce a Jacobi argument by half periods: u = 2K j + u_r, |u_r| <= K; returns j and the sign (-1)^j
KG_HD double reduce_half_periods(double u, double K, double i2K, double* ur, double* sgn) {
    double j = rint(u * i2K);               // i2K = 1/(2K), 0 when K = inf
    *ur = fma(-2.0 * K, j, u);
    *sgn = (((long long)j) & 1LL) ? -1.0 : 1.0;
    return j;
}

// (The hot track kernel does not come through here: kg_hot.cuh restates the theta branches below without the Carlson
// fallback, which takes 30-60 registers off that kernel.)
template <class W>
KG_HD void theta_eval(const W& w, int flags, double tau, bool full, double* th, double* Gph, double* Gt) {
    *th = 0.0; *Gph = 0.0; *Gt = 0.0;
    if (!(flags & (F_TH_ORD | F_TH_VORT))) return;
    double snarg = fma(-w(T_W), tau, w(T_X0));
    if (flags & F_TH_THETA) {   // theta-function path (kg_theta.cuh): no duplication, no period bookkeeping
        const double sq1m_ = w(T_SQ), ua = snarg * w(T_ISQ);
        double L = 0.0;
        const ThetaOut t = full ? theta_eval_motion<W, 1, 0x2>(w, T_TH, ua, &L) : theta_eval_motion<W, 0, 0>(w, T_TH, ua, &L);
        // sn/dn = (R1/R3) th1/th3, sn cn/dn = (R1 R2/R3) th1 th2/(th3 th4), Z = W th4'/th4: one reciprocal for all three
        const double inv34 = frcp(t.th3 * t.th4);
        const double sinA_ = sq1m_ * (w(T_TH + TB_R13) * t.th1 * (inv34 * t.th4));
        if (flags & F_TH_ORD) {
            *th = kg_acos(w(T_A) * sinA_);
        } else {
            double h = (flags & F_HPOS) ? 1.0 : -1.0;
            *th = kg_acos(h * sqrt(fma(w(T_B), sinA_ * sinA_, w(T_A))));
        }
        if (!full) return;
        *Gph = fma(w(T_TH_SLOPE), ua, w(T_TH_COEF) * L) - w(T_GPH0);
        const double Z_ = w(T_TH + TB_W) * t.dth4 * (inv34 * t.th3);
        const double Dtot = sq1m_ * (fma(w(T_TH_G), ua, Z_ * w(T_TH_IM)) - w(T_TH + TB_R123) * (t.th1 * t.th2) * inv34);
        *Gt = fma(w(T_C3), snarg, 3.0 * w(T_C1) * Dtot) - w(T_GT0);
        return;
    }
    double sq1m = w(T_SQ);
    const double uarg = snarg * w(T_ISQ);
    const SnCnDn jt = jacobi_from_ladder(SlotView<W>{w, T_LAD}, uarg);
    const double sn = jt.sn, cn = jt.cn, dn = jt.dn;
    const double Km = w(T_K);
    double idn = frcp(dn);
    double sinA = sq1m * sn * idn;            // sin(am(snarg|k)), kgeo/kerr_raytracing_ana.py:275
    if (flags & F_TH_ORD) {
        *th = kg_acos(w(T_A) * sinA);         // :280
    } else {
        double h = (flags & F_HPOS) ? 1.0 : -1.0;
        *th = kg_acos(h * sqrt(fma(w(T_B), sinA * sinA, w(T_A))));   // :355
    }
    if (!full) return;
    double ur, sg;
    double j = reduce_half_periods(uarg, Km, w(T_I2K), &ur, &sg);
    double s = sg * sinA, c = sg * cn * idn;
    double c2 = c * c, y = idn * idn;         // 1 - k sin^2 = 1/dn(u|m)^2
    double n = w(T_N);
    PiPrep pp = pi_prepare(n, s, c2, y);
    double rf, rd, rj;
    carlson_bundle<1, false, true>(c2, y, &pp.p, &rf, &rd, &rj);
    double Fr = ur * sq1m;                    // F(phi_r|k)
    double tail = pi_finish(n, s, c2, y, Fr, pp, rj);          // Pi - F on the reduced amplitude
    double gph = w(T_D3) * (snarg + tail);
    double gt = fma(w(T_C3), snarg, w(T_C1) * cube(s) * rd);
    if (j != 0.0) {
        gph = fma(w(T_D2), j, gph);           // +inf when n == 1 (kgeo/scipy_ellip_binding.py:149)
        gt = fma(w(T_C2), j, gt);
    }
    *Gph = gph - w(T_GPH0);
    *Gt = gt - w(T_GT0);
}

// Legendre E and three Pi's at the amplitude am(X|k) from the Jacobi functions at X
template <class W>
KG_HD void radial_legendre(const W& w, double X, double Kc, double sn, double cn, double dn, double* E, double* Pi) {
    double k = w(R_K);
    double Xr, sg;
    double j = reduce_half_periods(X, Kc, w(R_I2K), &Xr, &sg);
    double s = sg * sn, c2 = cn * cn, y = dn * dn;
    PiPrep pp[3];
    double pa[3], rj[3], rf, rd;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        pp[i] = pi_prepare(w(R_N0 + i), s, c2, y);
        pa[i] = pp[i].p;
    }
    carlson_bundle<3, false, true>(c2, y, pa, &rf, &rd, rj);
    double e = Xr - (k * (1.0 / 3.0)) * cube(s) * rd;
    if (j != 0.0) e = fma(2.0 * j, w(R_EC), e);
    *E = e;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double v = Xr + pi_finish(w(R_N0 + i), s, c2, y, Xr, pp[i], rj[i]);
        if (j != 0.0) v += 2.0 * j * (w(R_KC) + w(R_PC0 + i));
        Pi[i] = v;
    }
}

template <class W, bool LADDER = true>
KG_HD void radial_eval(const W& w, const Observer& o, int flags, double tau, bool full, double* r_s, double* Iph,
                       double* It, double* Isig) {
    double X = fma(-w(R_SC), tau, w(R_X0));
    double k = w(R_K);
    SnCnDn jr;
    double I1 = 0.0, I2 = 0.0, Ip = 0.0, Im = 0.0;
    // E and the three Pi's at am(X|k), shared by every radial case
    double E = 0.0, Pi[3] = {0.0, 0.0, 0.0};
    bool th_case2 = false;           // theta path of case 2: r(tau) and 1/den come out of the merged reciprocal below
    double rs2 = 0.0, iden2 = 0.0;
    if (LADDER && (flags & F_R_THETA)) {   // theta-function path (kg_theta.cuh)
        double L[3] = {0.0, 0.0, 0.0};
        ThetaOut t;
        if (!full) t = theta_eval_motion<W, 0, 0>(w, R_TH, X, L);
        else if (flags & F_R_K100) t = theta_eval_motion<W, 3, 0x001>(w, R_TH, X, L);
        else if (flags & F_R_K111) t = theta_eval_motion<W, 3, 0x111>(w, R_TH, X, L);
        else t = theta_eval_motion<W, 3, kKindsGeneric>(w, R_TH, X, L);
        double Z;
        if (flags & F_CASE2) {
            // r = (r4 r31 - r3 r41 sn^2)/(r31 - r41 sn^2) with sn = R1 th1/th4: numerator and denominator times th4^2, and
            // ONE reciprocal 1/(th4 D) gives both 1/th4 (for sn, cn, dn, Z) and 1/D
            th_case2 = true;
            const double q1 = w(R_Q0), q3 = w(R_Q2), q4 = w(R_Q3);
            const double rr31 = q3 - q1, rr41 = q4 - q1;
            const double s1 = w(R_TH + TB_R1) * t.th1, s2 = s1 * s1, t42 = t.th4 * t.th4;
            const double D = fma(-rr41, s2, rr31 * t42);
            const double inv = frcp(t.th4 * D);
            const double i4 = inv * D, iD = inv * t.th4;
            rs2 = fma(-(q3 * rr41), s2, (q4 * rr31) * t42) * iD;
            iden2 = t42 * iD;
            jr.sn = s1 * i4;
            jr.cn = w(R_TH + TB_R2) * t.th2 * i4;
            jr.dn = w(R_TH + TB_R3) * t.th3 * i4;
            Z = w(R_TH + TB_W) * t.dth4 * i4;
        } else {
            const ThetaJac j = theta_ratios(w, R_TH, t);
            jr.sn = j.sn; jr.cn = j.cn; jr.dn = j.dn;
            Z = j.Z;
        }
        if (full) {
            E = fma(w(R_TH + TB_EOKM1), X, X) + Z;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                Pi[i] = fma(w(R_TH + TB_SIZE + i * TP_SIZE + TP_A), X, w(R_TH + TB_SIZE + i * TP_SIZE + TP_P) * L[i]);
        }
    } else {
        if (LADDER) {
            jr = jacobi_from_ladder(SlotView<W>{w, R_LAD}, X);
        } else {
            double Kunused;
            jacobi_sncndn(X, k, &jr.sn, &jr.cn, &jr.dn, &Kunused);
        }
        // one shared Carlson bundle
        if (full) radial_legendre(w, X, w(R_KC), jr.sn, jr.cn, jr.dn, &E, Pi);
    }
    const double sn = jr.sn, cn = jr.cn, dn = jr.dn;
    const double rp = o.rplus, rm = o.rminus;
    if (flags & F_CASE2) {
        double q1 = w(R_Q0), q2 = w(R_Q1), q3 = w(R_Q2), q4 = w(R_Q3), root = w(R_Q4);
        double rr31 = q3 - q1, rr41 = q4 - q1, rr43 = q4 - q3;
        double rs, iden;
        if (th_case2) {
            rs = rs2;
            iden = iden2;
        } else {
            double sn2 = sn * sn;
            // explicit FMAs: r(tau) must not depend on how nvcc contracts the expression in the full / partial kernels
            double den = fma(-rr41, sn2, rr31);
            iden = frcp(den);
            rs = fma(-(q3 * rr41), sn2, q4 * rr31) * iden;                    // :592-594
        }
        *r_s = rs;
        if (full) {
            // H = (dr/dtau)/(r - r3) (:600-606) with dr/dtau = r31 r43 r41 sn cn dn root/den^2 and r - r3 = r31 r43/den
            double H = rr41 * (sn * cn * dn * root) * iden - w(R_Q5);
            double Pi_1 = w(R_Q6) * (Pi[0] - w(R_PO0));
            double Pi_p = w(R_Q7) * (Pi[1] - w(R_PO1));
            double Pi_m = w(R_Q8) * (Pi[2] - w(R_PO2));
            I1 = q3 * (-tau) + rr43 * Pi_1;
            I2 = H - 0.5 * (q1 * q4 + q2 * q3) * (-tau) - root * (E - w(R_EO));
            Ip = tau * w(R_Q9) - Pi_p;
            Im = tau * w(R_Q10) - Pi_m;
        }
    } else if (flags & F_CASE3) {
        double rr1 = w(R_Q0), rr2 = w(R_Q1), AA = w(R_Q2), BB = w(R_Q3), sqAB = w(R_Q10);
        const double b2 = BB * rr2, a1_ = AA * rr1;
        *r_s = fdiv(fma(b2 + a1_, cn, b2 - a1_), fma(BB + AA, cn, BB - AA));   // :516-519
        if (full) {
            // R1 = (Pi - al f1)/(1 - al^2), f1 = (p1/2) ln|(p1 dn + sn)/(p1 dn - sn)| (B62; :666-674); reciprocals per ray
            double R1a[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double al = w(R_Q4 + i), p1 = w(R_Q7 + i);
                const double f1 = p1 * atanh_ratio(p1 * dn, sn);
                R1a[i] = (Pi[i] - al * f1) * w(R_C0 + i);
            }
            // R2 (B65; :675-683)
            const double al0 = w(R_Q4), al2 = al0 * al0, iden = w(R_C9);
            double R2a = (X - (al2 * iden) * (E - fdiv(al0 * sn * dn, 1.0 + al0 * cn))) * w(R_C10);
            R2a += (2.0 * k - w(R_N0)) * R1a[0] * iden;
            double dm = (flags & F_LOWSPIN) ? 0.0 : (R1a[2] - w(R_PO2));      // :529-534
            double f0 = w(R_C3);
            double Pi_1 = f0 * (R1a[0] - w(R_PO0));
            double Pi_2 = sq(f0) * (R2a - w(R_EO));
            double Pi_p = w(R_C4) * (R1a[1] - w(R_PO1));
            double Pi_m = w(R_C5) * dm;
            double pr = w(R_C6), bpa = BB + AA;
            I1 = pr * (-tau) + Pi_1;
            I2 = pr * pr * (-tau) + 2.0 * pr * Pi_1 + sqAB * Pi_2;
            Ip = (bpa * (-tau) + Pi_p) * w(R_C7);
            Im = (bpa * (-tau) + Pi_m) * w(R_C8);
        }
    } else {  // case 4
        double a2 = w(R_Q0), b1 = w(R_Q1), g0 = w(R_Q2);
        *r_s = fma(-a2, fdiv(fma(g0, cn, -sn), fma(g0, sn, cn)), -b1);        // :441 with sc = sn/cn
        if (full) {
            // S1 = (F + al^2 Pi - al f2)/(1 + al^2), f2 = (p2/2) ln|(1-p2)(1+p2 dn)/((1+p2)(1-p2 dn))| (B92; :638-650): the
            // dn-independent half of the logarithm is a per-ray constant
            double S1a[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double al = w(R_Q2 + i), p2 = w(R_Q5 + i);
                const double f2 = p2 * (atanh_ratio(1.0, p2 * dn) + w(R_C3 + i));
                S1a[i] = (X + al * al * Pi[i] - al * f2) * w(R_C0 + i);
            }
            // S2 (B93; :651-663)
            const double al2 = g0 * g0, i1 = w(R_C0), i2 = w(R_C11);
            double S2a = (1.0 - k) * X + al2 * E + fdiv(al2 * dn * (g0 * cn - sn), cn + g0 * sn) - al2 * g0;
            S2a = -S2a * (i1 * i2) + S1a[0] * (i1 + (1.0 - k) * i2);
            double pref = w(R_Q8);
            double a2g = w(R_C6);
            double aa0 = a2g * (1.0 + al2);
            double Pi_1 = pref * aa0 * (S1a[0] - w(R_PO0));
            double Pi_2 = pref * (aa0 * aa0) * (S2a - w(R_EO));
            double Pi_p = w(R_C7) * (S1a[1] - w(R_PO1));
            double Pi_m = w(R_C8) * (S1a[2] - w(R_PO2));
            double aa1 = a2g - b1;
            I1 = aa1 * (-tau) - Pi_1;
            I2 = (aa1 * aa1) * (-tau) - 2.0 * aa1 * Pi_1 + Pi_2;
            Ip = w(R_C9) * (-tau - Pi_p);
            Im = w(R_C10) * (-tau - Pi_m);
        }
    }
    if (!full) {
        *Iph = 0.0; *It = 0.0; *Isig = 0.0;
        return;
    }
    // kgeo/kerr_raytracing_ana.py:626-631
    double a = o.a, lam = w(S_LAM);
    *Iph = o.c_iph * ((rp - 0.5 * a * lam) * Ip - (rm - 0.5 * a * lam) * Im);
    double ItA = o.c_it * ((rp * rp - 0.5 * a * lam * rp) * Ip - (rm * rm - 0.5 * a * lam * rm) * Im);
    *It = ItA + 4.0 * (-tau) + 2.0 * I1 + I2;
    *Isig = I2;
}

// one full point (kgeo/kerr_raytracing_ana.py:172-188)
template <class W>
KG_HD Point point_eval(const W& w, const Observer& o, int flags, double tau, bool full) {
    Point p;
    double Gph, Gt, Iph, It, Isig;
    theta_eval<W>(w, flags, tau, full, &p.th, &Gph, &Gt);
    radial_eval<W, true>(w, o, flags, tau, full, &p.r, &Iph, &It, &Isig);
    p.sig = Isig + o.a2 * Gt;
    p.t = It + o.a2 * Gt;
    p.ph = Iph + w(S_LAM) * Gph;
    return p;
}

// ---- crossing counts -------------------------------------------------------------------------------------------
// n_equatorial_crossings (utils :605-684) and n_poloidal_orbits (:551-603) at Mino time tau
template <class W>
KG_HD double n_equatorial_at(const W& w, const Observer& o, int flags, double beta, double tau) {
    if (!(w(S_ETA) > 0.0)) return -2.0;
    double F0 = w(T_F0), K = w(T_KC), ww = w(T_W);
    bool first = o.north ? (beta <= 0.0) : (beta >= 0.0);
    double sgn = (o.north == (first ? 1 : 0)) ? -1.0 : 1.0;   // north&first: -F0 ; north&!first: +F0 ; south: mirrored
    double v = floor((tau * ww + sgn * F0) / (2.0 * K));
    return first ? v : v - 1.0;
}
template <class W>
KG_HD double n_poloidal_at(const W& w, double tau) {
    if (!(w(S_ETA) > 0.0)) return -2.0;
    return (w(T_W) * tau) / (4.0 * w(T_KC));
}

// accessor over a plain array (registers / local)
struct ArrayW {
    const double* p;
    KG_HD double operator()(int i) const { return p[i]; }
};
// accessor over the grouped workspace in HBM: slot i of ray `pix` lives at w[ws_offset(pix) + i * kRaysPerGroup]
struct StridedW {
    const double* base;      // w + ws_offset(pix)
    KG_HD double operator()(int i) const { return base[i * kRaysPerGroup]; }
};
KG_HD StridedW ray_view(const double* w, size_t pix) { return StridedW{w + ws_offset(pix)}; }
KG_HD WsPtr ray_ptr(double* w, size_t pix) { return WsPtr{w + ws_offset(pix), (size_t)kRaysPerGroup}; }

}  // namespace kg
