// Theta-function evaluation of sn/cn/dn, E and Pi along a ray (sm_100a).
//
// Why: along one ray the Jacobi parameter m and the characteristics n_i of the elliptic integrals are CONSTANTS; only the
// Mino phase u changes from sample to sample.  Carlson duplication (the SciPy / GSL algorithm the reference uses through
// kgeo/scipy_ellip_binding.py:87-237) re-derives everything from scratch for every sample: ~4.5 duplication steps, each
// with 3 + NP square roots, NP reciprocals and NP short series.  With the nome q = exp(-pi K'/K) of the ray the same
// functions are Fourier-type series in v = pi u / 2K whose coefficients are per-ray constants:
//
//   sn u = [th3(0)/th2(0)] th1(v)/th4(v),  cn u = [th4(0)/th2(0)] th2(v)/th4(v),  dn u = [th4(0)/th3(0)] th3(v)/th4(v)
//   E(am u|m) = (E/K) u + Z(u),   Z(u) = (pi/2K) th4'(v)/th4(v)
//   Pi(n; am u|m) = u + P(a) [ u Z(a) + (1/2) ln( th4(v - va) / th4(v + va) ) ],   sn^2(a|m) = n/m,  P = sn a/(cn a dn a)
//
// (Jacobi's form of the third-kind integral).  The constant a is real for 0 < n < m, b + iK' for n > 1 (the logarithm turns
// into ln|th1(v - vb)/th1(v + vb)|: exactly the Cauchy principal value the reference takes for 1 - n sin^2 phi < 0,
// scipy_ellip_binding.py:207-230), K + ib for m < n < 1 and ib for n < 0 (the logarithm turns into the argument of
// th3 resp. th4 at v - i beta).  All four forms were checked against 40-digit quadrature in research/theta_pi_*.py
// (<= 5e-15 relative in float64).  The series converge like q^(j^2): over BASELINE cfg 3 the median nome is 8e-4 (polar)
// and 0.017 (radial), i.e. 2-4 terms; rays with q > 0.3 or an ill-conditioned prefactor (next to the critical curve,
// n -> m, n -> 1) keep the Carlson path, decided once per ray in the setup.
#pragma once
#include "kg_theta_layout.cuh"

namespace kg {

// F(psi|m) and D(psi|m) = (F - E)/m = s^3 RD / 3 for an amplitude given by (sin, cos) in the first quadrant
KG_NOINLINE void legendre_F_D(double s, double c, double m, double* F, double* D) {
    double c2 = c * c, y = fma(-m * s, s, 1.0);
    double rf, rd;
    carlson_bundle<0, true, true>(c2, y, nullptr, &rf, &rd, nullptr);
    *F = s * rf;
    *D = cube(s) * rd * (1.0 / 3.0);
}

// Rays outside these bounds keep the Carlson path (decided once per ray in the setup):
//   nome: the series need J = 7 harmonics at q = 0.4 (kThJM = 8 are stored); beyond that the ray is within ~1e-3 of the
//   critical curve, where the reference itself loses digits (looser gate next to the critical curve);
//   prefactor P = sn a/(cn a dn a) of Pi = A u + P L: A = 1 + P Z(a) is formed with an absolute error ~ P eps (radial bound
//   3000, and n no closer than 1e-6 to m or 1);
//   for the polar motion (kind 2, L = arg theta3(v - i beta)) a ray that passes next to the rotation axis drives theta3 to
//   within ~1/P of its zero, so L carries eps P and Pi eps P^2 there (measured 9e-11 at P = 780) -- but ONLY there, within
//   ~1/P of the pole, where the azimuth is a coordinate singularity and phi is gated on the arc length sin(theta) dphi
//   (the "axis" gate of the parity tests: the arc-length error stays ~eps P).  Polar bound 4000 / edge 1e-8: the two pixel columns next to
//   alpha = 0 of a 1024-wide image (2/3 of all fallback rays of BASELINE cfg 3 under the earlier bound of 300) take the hot pass.
#ifndef KG_THETA_QMAX
#define KG_THETA_QMAX 0.4
#endif
#ifndef KG_THETA_PMAX
#define KG_THETA_PMAX 3000.0
#endif
#ifndef KG_THETA_PMAX_POLAR
#define KG_THETA_PMAX_POLAR 4000.0
#endif
#ifndef KG_THETA_EDGE
#define KG_THETA_EDGE 1.0e-6          // n within this (relative) distance of m or 1: no theta form (radial motion)
#endif
#ifndef KG_THETA_EDGE_POLAR
#define KG_THETA_EDGE_POLAR 1.0e-8
#endif
constexpr double kThetaQMax = KG_THETA_QMAX, kThetaPMax = KG_THETA_PMAX, kThetaPMaxPolar = KG_THETA_PMAX_POLAR;
constexpr double kThetaEdge = KG_THETA_EDGE, kThetaEdgePolar = KG_THETA_EDGE_POLAR;

struct ThetaRay {
    double m, K, Kp, q, eokm1, eokm1p;   // eokm1p = E(m')/K' - 1
    double RDc;                          // RD(0, 1 - m, 1) = 3 (K - E) / m
    int ok;
};

// K_known / RDc_known: K(m) and RD(0, 1-m, 1) when the caller already has them (the first setup kernel stores both: every
// complete integral is evaluated once per ray), NaN to have them computed here
KG_HD ThetaRay theta_ray(double m, double K_known = kInf - kInf, double RDc_known = kInf - kInf) {
    ThetaRay r;
    r.m = m;
    r.ok = (m > 1.0e-12 && m < 1.0) ? 1 : 0;
    if (!r.ok) {
        r.K = r.Kp = r.q = r.eokm1 = r.eokm1p = r.RDc = 0.0;
        return r;
    }
    double mp = 1.0 - m;
    const bool haveK = (K_known == K_known), haveRD = (RDc_known == RDc_known);
    r.K = K_known;
    r.RDc = RDc_known;
    if (!haveK || !haveRD) {
        double Kc, RDm;
        ellip_complete<0>(m, nullptr, &Kc, &RDm, nullptr);
        if (!haveK) r.K = Kc;
        if (!haveRD) r.RDc = RDm;
    }
    double RDp;
    ellip_complete<0>(mp, nullptr, &r.Kp, &RDp, nullptr);               // complementary parameter: K', RD(0, m, 1)
    r.q = exp(-kPi * r.Kp / r.K);
    r.eokm1 = -m * r.RDc / (3.0 * r.K);
    r.eokm1p = -mp * RDp / (3.0 * r.Kp);
    if (!(r.q <= kThetaQMax)) r.ok = 0;
    return r;
}

// Powers of the nome by recurrence: q^((j+1)^2) = q^(j^2) q^(2j+1), q^((j+1/2)^2) = q^(1/4) q^(j^2+j); the relative error
// grows like j eps, irrelevant for coefficients that are themselves < q^(j^2).
struct NomePowers {
    double q2, qe, qo, stepe, stepo;   // qe = q^((j+1)^2), qo = q^((j+1/2)^2) for the current j
    KG_HD explicit NomePowers(double q) : q2(q * q), qe(q), qo(sqrt(sqrt(q))), stepe(q * q * q), stepo(q * q) {}
    KG_HD void next() {
        qe *= stepe; stepe *= q2;      // q^((j+2)^2) = q^((j+1)^2) q^(2j+3)
        qo *= stepo; stepo *= q2;      // q^((j+3/2)^2) = q^((j+1/2)^2) q^(2j+2)
    }
};

// base block of a motion; returns the number of harmonics needed by the base series
template <class P>
KG_HD int theta_make_base(const ThetaRay& r, P out) {
    for (int i = 0; i < TB_SIZE; ++i) out[i] = 0.0;
    if (!r.ok) return 0;
    out[TB_W] = kHalfPi / r.K;
    out[TB_EOKM1] = r.eokm1;
    double t2 = 0.0, t3 = 1.0, t4 = 1.0, sg = 1.0;
    int J = kThJM;
    NomePowers np(r.q);
    for (int j = 0; j < kThJM; ++j) {
        const double qe = np.qe, qo = np.qo;
        if (qe < 1.0e-17 && qo < 1.0e-17) {     // harmonics 1..j (even) / 0..j-1 (odd) suffice
            J = j;
            break;
        }
        out[TB_QO + j] = 2.0 * sg * qo;      // (-1)^j
        out[TB_QE + j] = -2.0 * sg * qe;     // (-1)^(j+1)
        t2 += 2.0 * qo;
        t3 += 2.0 * qe;
        t4 += -2.0 * sg * qe;
        sg = -sg;
        np.next();
    }
    if (J < 1) J = 1;
    out[TB_R1] = t3 / t2;
    out[TB_R2] = t4 / t2;
    out[TB_R3] = t4 / t3;
    out[TB_R13] = (t3 * t3) / (t2 * t4);
    out[TB_R123] = (t3 * t3) / (t2 * t2);
    return J;
}

// one Pi block; returns the number of harmonics it needs, or -1 when the theta form is unusable for this (n, m)
// nmm = n - m and omn = 1 - n are passed separately so that callers who know them without cancellation (the polar motion:
// n' - m = n (1 - m), 1 - n' = (1 - n)(1 - m)) keep full accuracy next to the edges n -> m, n -> 1.
template <class WP>
KG_HD int theta_make_pi(const ThetaRay& r, double n, double nmm, double omn, double pmax, double edge, WP out) {
    for (int i = 0; i < TP_SIZE; ++i) out[i] = 0.0;
    out[TP_A] = 1.0;
    if (!r.ok || !(n == n)) return -1;
    const double m = r.m, mp = 1.0 - m, q = r.q;
    if (n == 0.0) return 1;                                    // Pi = F = u
    if (fabs(nmm) < edge * m || fabs(omn) < edge) return -1;
    double sh, P, Zs;
    int kind;
    if (n > 0.0 && n < m) {              // a real: sn a = sqrt(n/m)
        double s = sqrt(n / m), c = sqrt(-nmm / m), d = sqrt(omn);
        double F, D;
        legendre_F_D(s, c, m, &F, &D);
        Zs = -m * D - r.eokm1 * F;                            // Z(a) = E(a) - (E/K) a
        P = s / (c * d);
        sh = kHalfPi * F / r.K;
        kind = 0;
    } else if (n > 1.0) {                // a = b + iK': sn b = 1/sqrt(n)
        double s = sqrt(1.0 / n), c = sqrt(-omn / n), d = sqrt(nmm / n);
        double F, D;
        legendre_F_D(s, c, m, &F, &D);
        Zs = (-m * D - r.eokm1 * F) + c * d / s;              // Z(b) + cs b dn b
        P = -s / (c * d);
        sh = kHalfPi * F / r.K;
        kind = 1;
    } else {                             // imaginary shift, complementary modulus
        double s, c, d;
        if (n < 0.0) {                   // a = ib: sc^2(b|m') = -n/m
            double t = sqrt(-n / m);
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
        } else {                         // a = K + ib: dn^2(b|m') = m/n
            s = sqrt(nmm / (n * mp));
            c = sqrt(m * omn / (n * mp));
        }
        d = (n < 0.0) ? sqrt(fma(-mp * s, s, 1.0)) : sqrt(m / n);
        double F, D;
        legendre_F_D(s, c, mp, &F, &D);
        double Zbp = -mp * D - r.eokm1p * F;                  // Z(b|m')
        double Zd = d * s / c - Zbp - kPi * F / (2.0 * r.K * r.Kp);
        if (n < 0.0) {
            P = -(s * c / d);
            Zs = Zd;
        } else {
            P = -(d / (mp * s * c));
            Zs = Zd - m * s / (c * d);
        }
        sh = kHalfPi * F / r.K;
        kind = 2;
    }
    if (!(fabs(P) <= pmax) || !(Zs == Zs)) return -1;
    out[TP_KIND] = (double)kind;
    out[TP_P] = P;
    out[TP_A] = fma(P, Zs, 1.0);
    // harmonics of the shift by rotation (real shift) or by powers of e^(2 sh) (imaginary shift)
    int J = -1;
    double sg = 1.0;
    NomePowers np(q);
    double c1, s1, cj, sj, ep = 1.0, em = 1.0, e2p = 1.0, e2m = 1.0;
    if (kind == 2) {
        e2p = exp(2.0 * sh);
        e2m = 1.0 / e2p;
        ep = e2p; em = e2m;              // e^(+-2 (j+1) sh)
        c1 = s1 = cj = sj = 0.0;
    } else {
        double ss = sin(sh), cs = cos(sh);
        c1 = fma(cs, cs, -ss * ss); s1 = 2.0 * ss * cs;      // cos 2 sh, sin 2 sh
        if (kind == 0) { cj = c1; sj = s1; }                   // (cos, sin)(2 (j+1) sh)
        else { cj = cs; sj = ss; }                             // (cos, sin)((2j+1) sh)
    }
    for (int j = 0; j < kThJM; ++j) {
        double cx, sy, mag;
        if (kind == 0) {                 // 2 (-1)^(j+1) q^((j+1)^2) {cos, sin}(2 (j+1) sh)
            double qe = -2.0 * sg * np.qe;
            cx = qe * cj;
            sy = qe * sj;
            mag = fabs(qe);
        } else if (kind == 1) {          // 2 (-1)^j q^((j+1/2)^2) {cos, sin}((2j+1) sh); th1(v -+ sh) = X +- Y:
            double qo = 2.0 * sg * np.qo; //   L = (1/2) ln|th1(v-sh)/th1(v+sh)| = (1/2) ln|(X+Y)/(X-Y)|
            cx = qo * cj;
            sy = -qo * sj;
            mag = fabs(qo);
        } else {                         // th4 (n<0): 2 (-1)^(j+1) q^(..) {cosh, sinh};  th3 (m<n<1): 2 q^(..) {cosh, sinh}
            double qe = 2.0 * np.qe * ((n < 0.0) ? -sg : 1.0);
            cx = qe * (0.5 * (ep + em));
            sy = qe * (0.5 * (ep - em));
            mag = fabs(cx);
        }
        if (mag < 1.0e-17) {
            J = j;
            break;
        }
        out[TP_C + j] = cx;
        out[TP_S + j] = sy;
        sg = -sg;
        np.next();
        if (kind == 2) {
            ep *= e2p; em *= e2m;
        } else {
            double cn_ = fma(cj, c1, -sj * s1), sn_ = fma(sj, c1, cj * s1);
            cj = cn_; sj = sn_;
        }
    }
    if (J < 0) {
        // not converged within the stored harmonics
        double last = fabs(out[TP_C + kThJM - 1]) + fabs(out[TP_S + kThJM - 1]);
        if (last > 1.0e-17) return -1;
        J = kThJM;
    }
    if (J < 1) J = 1;
    return J;
}

// ---- per-sample evaluation ------------------------------------------------------------------------------------------
// Out: the theta sums (-> sn, cn, dn, Z(u) by theta_ratios) and L_i (i < NP) of the motion whose base block starts at slot `base`; Pi blocks follow the base block.
struct ThetaOut {
    double th1, th2, th3, th4, dth4;   // theta1..4(v) (up to the constant factors folded into TB_R*) and d theta4/dv
};
// sn, cn, dn, Z from the raw sums with one reciprocal (callers that need further quotients merge their own, see
// theta_eval / radial_eval)
struct ThetaJac {
    double sn, cn, dn, Z;
};
template <class W>
KG_HD ThetaJac theta_ratios(const W& w, int base, const ThetaOut& t) {
    const double i4 = frcp(t.th4);
    return ThetaJac{w(base + TB_R1) * t.th1 * i4, w(base + TB_R2) * t.th2 * i4, w(base + TB_R3) * t.th3 * i4,
                    w(base + TB_W) * t.dth4 * i4};
}

// KINDS packs the kind (0, 1 or 2, see ThetaPiSlot) of Pi block i into bits [4i, 4i+3]; a nibble of 0xF means "read it from
// the block" (generic).  The setup flags the rays whose pattern is one of the static ones (the polar block is always kind
// 2; radial (1,0,0) = rays that turn outside the horizon, (1,1,1) = plunging rays), so the harmonic loop of almost every
// warp has no selects.  The loop is fully unrolled (J <= kThJM, early exit): slot offsets become immediates and the
// rotated harmonics stay in place.
constexpr int kKindsGeneric = 0xFFF;
template <class W, int NP, int KINDS>
KG_HD ThetaOut theta_eval_motion(const W& w, int base, double u, double* L) {
    // J >= 1 (theta_block_build): saying so lets the compiler drop the exit in front of the first harmonic, and with it the
    // 11-16 instructions that would materialise the initial values of the sums (they fold into the first FMAs instead)
    const int Jw = (int)w(base + TB_J), J = (Jw < 1) ? 1 : Jw;
    const double v = w(base + TB_W) * u;
    double sv, cv;
    fsincos(v, &sv, &cv);
    const double c1 = fma(cv, cv, -sv * sv), s1 = 2.0 * sv * cv;   // cos 2v, sin 2v
    // even harmonics (ce, se) = (cos, sin)(2(j+1)v) and odd ones (co, so) = (cos, sin)((2j+1)v) by the three-term recurrence
    // f_(j+1) = 2 cos(2v) f_j - f_(j-1): one FMA per sequence and harmonic (error ~ j^2 eps on terms that are < q^(j^2 - j))
    const double tc = c1 + c1;
    double ce = c1, se = s1, cep = 1.0, sep = 0.0;       // j = 0: 2v; previous: 0
    double co = cv, so = sv, cop = cv, sop = -sv;        // j = 0: v;  previous: -v
    double th1 = 0.0, th2 = 0.0, th3 = 1.0, th4 = 1.0, dth4 = 0.0;
    double X[NP > 0 ? NP : 1], Y[NP > 0 ? NP : 1];
    int kind[NP > 0 ? NP : 1];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int ks = (KINDS >> (4 * i)) & 0xF;
        kind[i] = (ks == 0xF) ? (int)w(base + TB_SIZE + i * TP_SIZE + TP_KIND) : ks;
        X[i] = (kind[i] == 1) ? 0.0 : 1.0;
        Y[i] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < kThJM; ++j) {
        if (j >= J) break;
        const double qe = w(base + TB_QE + j), qo = w(base + TB_QO + j);   // 2(-1)^(j+1) q^((j+1)^2), 2(-1)^j q^((j+1/2)^2)
        const double sg = (j & 1) ? -1.0 : 1.0;
        th4 = fma(qe, ce, th4);
        th3 = fma(-sg * qe, ce, th3);            // 2 q^((j+1)^2) cos
        dth4 = fma(-2.0 * (double)(j + 1) * qe, se, dth4);
        th1 = fma(qo, so, th1);
        th2 = fma(sg * qo, co, th2);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int pb = base + TB_SIZE + i * TP_SIZE;
            const double hc = (kind[i] == 1) ? so : ce;
            const double hs = (kind[i] == 1) ? co : se;
            X[i] = fma(w(pb + TP_C + j), hc, X[i]);
            Y[i] = fma(w(pb + TP_S + j), hs, Y[i]);
        }
        // next harmonics
        const double cen = fma(tc, ce, -cep), sen = fma(tc, se, -sep);
        const double con = fma(tc, co, -cop), son = fma(tc, so, -sop);
        cep = ce; sep = se; cop = co; sop = so;
        ce = cen; se = sen; co = con; so = son;
    }
    const ThetaOut o{th1, th2, th3, th4, dth4};
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        if (kind[i] == 2) L[i] = fatan2(Y[i], X[i]);
        else L[i] = atanh_ratio(X[i], Y[i]);
    }
    return o;
}

}  // namespace kg
