// The hot per-sample code of a full (r, theta, phi, t) trace: theta-function path only, NS Mino-time samples of ONE ray per
// thread.
//
// Same formulas, in the same operation order, as the theta branches of theta_eval / radial_eval (kg_geodesic.cuh; those stay
// the complete code: Carlson fallback, partial outputs, image / lensing callers).  What changes is the shape of the work:
//  * every ray constant is read from shared memory ONCE per NS samples (the r1 kernel issued 92 shared-memory loads per
//    point; they were 11 % of its instructions and its LSU data pipe was 73 % busy), and so is every branch on a ray
//    property (radial case, Pi kinds, harmonic count, polar class);
//  * the NS samples are independent dependency chains in one thread: the harmonic recurrences, the reciprocals and the case
//    algebra of sample 0 fill the FP64-pipe latency of sample 1 (ILP instead of resident warps).
// r and theta are bit-identical to the complete code (explicit FMAs on that path); phi and t agree to the last bits (nvcc
// may contract their final sums differently).
//
// Reference formulas: kgeo/kerr_raytracing_ana.py:195-365 (th_integrate), :368-633 (r_integrate), :638-683 (S1_S2, R1_R2).
#pragma once
#include "kg_geodesic.cuh"

namespace kg {

#define KG_S_LOOP _Pragma("unroll") for (int s = 0; s < NS; ++s)

// theta sums of NS samples of one motion (see theta_eval_motion); L[s][i] = the Pi logarithm / angle of Pi block i
template <class W, int NP, int KINDS, int NS>
KG_HD void theta_motion_ns(const W& w, int base, const double (&u)[NS], ThetaOut (&out)[NS], double (&L)[NS][3]) {
    constexpr int NPA = NP > 0 ? NP : 1;
    const int Jw = (int)w(base + TB_J), J = (Jw < 1) ? 1 : Jw;
    const double wv = w(base + TB_W);
    double tc[NS], ce[NS], se[NS], cep[NS], sep[NS], co[NS], so[NS], cop[NS], sop[NS];
    double th1[NS], th2[NS], th3[NS], th4[NS], dth4[NS];
    double X[NS][NPA], Y[NS][NPA];
    int kind[NPA];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int ks = (KINDS >> (4 * i)) & 0xF;
        kind[i] = (ks == 0xF) ? (int)w(base + TB_SIZE + i * TP_SIZE + TP_KIND) : ks;
    }
    KG_S_LOOP {
        const double v = wv * u[s];
        double sv, cv;
        fsincos(v, &sv, &cv);
        const double c1 = fma(cv, cv, -sv * sv), s1 = 2.0 * sv * cv;   // cos 2v, sin 2v
        tc[s] = c1 + c1;
        ce[s] = c1; se[s] = s1; cep[s] = 1.0; sep[s] = 0.0;            // even harmonics: j = 0 is 2v, previous 0
        co[s] = cv; so[s] = sv; cop[s] = cv; sop[s] = -sv;             // odd harmonics:  j = 0 is v,  previous -v
        th1[s] = 0.0; th2[s] = 0.0; th3[s] = 1.0; th4[s] = 1.0; dth4[s] = 0.0;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            X[s][i] = (kind[i] == 1) ? 0.0 : 1.0;
            Y[s][i] = 0.0;
        }
    }
#pragma unroll
    for (int j = 0; j < kThJM; ++j) {
        if (j >= J) break;
        if (j > 0) {      // harmonics j from j - 1 and j - 2 (advancing at the top of the step: none is computed in vain)
            KG_S_LOOP {
                const double cen = fma(tc[s], ce[s], -cep[s]), sen = fma(tc[s], se[s], -sep[s]);
                const double con = fma(tc[s], co[s], -cop[s]), son = fma(tc[s], so[s], -sop[s]);
                cep[s] = ce[s]; sep[s] = se[s]; cop[s] = co[s]; sop[s] = so[s];
                ce[s] = cen; se[s] = sen; co[s] = con; so[s] = son;
            }
        }
        const double qe = w(base + TB_QE + j), qo = w(base + TB_QO + j);
        double pc[NPA], ps[NPA];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int pb = base + TB_SIZE + i * TP_SIZE;
            pc[i] = w(pb + TP_C + j);
            ps[i] = w(pb + TP_S + j);
        }
        const double sg = (j & 1) ? -1.0 : 1.0;
        const double q3 = -sg * qe, qd = -2.0 * (double)(j + 1) * qe, q2 = sg * qo;
        KG_S_LOOP {
            th4[s] = fma(qe, ce[s], th4[s]);
            th3[s] = fma(q3, ce[s], th3[s]);
            dth4[s] = fma(qd, se[s], dth4[s]);
            th1[s] = fma(qo, so[s], th1[s]);
            th2[s] = fma(q2, co[s], th2[s]);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const double hc = (kind[i] == 1) ? so[s] : ce[s];
                const double hs = (kind[i] == 1) ? co[s] : se[s];
                X[s][i] = fma(pc[i], hc, X[s][i]);
                Y[s][i] = fma(ps[i], hs, Y[s][i]);
            }
        }
    }
    KG_S_LOOP {
        out[s] = ThetaOut{th1[s], th2[s], th3[s], th4[s], dth4[s]};
#pragma unroll
        for (int i = 0; i < NP; ++i) L[s][i] = (kind[i] == 2) ? fatan2(Y[s][i], X[s][i]) : atanh_ratio(X[s][i], Y[s][i]);
    }
}

// polar motion: theta, G_phi, G_t of NS samples (theta branch of theta_eval, full = true)
template <class W, int NS>
KG_HD void polar_hot(const W& w, int flags, const double (&tau)[NS], double (&th)[NS], double (&Gph)[NS], double (&Gt)[NS]) {
    if (!(flags & (F_TH_ORD | F_TH_VORT))) {
        KG_S_LOOP { th[s] = 0.0; Gph[s] = 0.0; Gt[s] = 0.0; }
        return;
    }
    const double tw = w(T_W), x0 = w(T_X0), sq1m = w(T_SQ), isq = w(T_ISQ);
    double snarg[NS], ua[NS];
    KG_S_LOOP {
        snarg[s] = fma(-tw, tau[s], x0);
        ua[s] = snarg[s] * isq;
    }
    ThetaOut t[NS];
    double L[NS][3];
    theta_motion_ns<W, 1, 0x2, NS>(w, T_TH, ua, t, L);
    const double r13 = w(T_TH + TB_R13), tA = w(T_A);
    double inv34[NS];
    if (flags & F_TH_ORD) {
        KG_S_LOOP {
            inv34[s] = frcp(t[s].th3 * t[s].th4);
            const double sinA = sq1m * (r13 * t[s].th1 * (inv34[s] * t[s].th4));
            th[s] = kg_acos(tA * sinA);
        }
    } else {
        const double h = (flags & F_HPOS) ? 1.0 : -1.0, tB = w(T_B);
        KG_S_LOOP {
            inv34[s] = frcp(t[s].th3 * t[s].th4);
            const double sinA = sq1m * (r13 * t[s].th1 * (inv34[s] * t[s].th4));
            th[s] = kg_acos(h * sqrt(fma(tB, sinA * sinA, tA)));
        }
    }
    const double slope = w(T_TH_SLOPE), coef = w(T_TH_COEF), gph0 = w(T_GPH0);
    const double bw = w(T_TH + TB_W), g = w(T_TH_G), im = w(T_TH_IM), r123 = w(T_TH + TB_R123);
    const double c3 = w(T_C3), c1x3 = 3.0 * w(T_C1), gt0 = w(T_GT0);
    KG_S_LOOP {
        Gph[s] = fma(slope, ua[s], coef * L[s][0]) - gph0;
        const double Z = bw * t[s].dth4 * (inv34[s] * t[s].th3);
        const double Dtot = sq1m * (fma(g, ua[s], Z * im) - r123 * (t[s].th1 * t[s].th2) * inv34[s]);
        Gt[s] = fma(c3, snarg[s], c1x3 * Dtot) - gt0;
    }
}

// radial motion: r, I_phi, I_t, I_sigma of NS samples (theta branch of radial_eval, full = true)
template <class W, int NS>
KG_HD void radial_hot(const W& w, const Observer& o, int flags, const double (&tau)[NS], double (&rs)[NS], double (&Iph)[NS],
                      double (&It)[NS], double (&Isig)[NS]) {
    const double sc = w(R_SC), x0 = w(R_X0), k = w(R_K);
    double X[NS];
    KG_S_LOOP X[s] = fma(-sc, tau[s], x0);
    ThetaOut t[NS];
    double L[NS][3];
    if (flags & F_R_K100) theta_motion_ns<W, 3, 0x001, NS>(w, R_TH, X, t, L);
    else if (flags & F_R_K111) theta_motion_ns<W, 3, 0x111, NS>(w, R_TH, X, t, L);
    else theta_motion_ns<W, 3, kKindsGeneric, NS>(w, R_TH, X, t, L);
    double sn[NS], cn[NS], dn[NS], Z[NS], iden2[NS];
    const double br1 = w(R_TH + TB_R1), br2 = w(R_TH + TB_R2), br3 = w(R_TH + TB_R3), bw = w(R_TH + TB_W);
    if (flags & F_CASE2) {
        // r = (r4 r31 - r3 r41 sn^2)/(r31 - r41 sn^2), sn = R1 th1/th4: one reciprocal 1/(th4 D) gives 1/th4 and 1/D
        const double q1 = w(R_Q0), q3 = w(R_Q2), q4 = w(R_Q3);
        const double rr31 = q3 - q1, rr41 = q4 - q1;
        const double n3 = -(q3 * rr41), n4 = q4 * rr31;
        KG_S_LOOP {
            const double s1 = br1 * t[s].th1, s2 = s1 * s1, t42 = t[s].th4 * t[s].th4;
            const double D = fma(-rr41, s2, rr31 * t42);
            const double inv = frcp(t[s].th4 * D);
            const double i4 = inv * D, iD = inv * t[s].th4;
            rs[s] = fma(n3, s2, n4 * t42) * iD;
            iden2[s] = t42 * iD;
            sn[s] = s1 * i4;
            cn[s] = br2 * t[s].th2 * i4;
            dn[s] = br3 * t[s].th3 * i4;
            Z[s] = bw * t[s].dth4 * i4;
        }
    } else {
        KG_S_LOOP {
            const double i4 = frcp(t[s].th4);
            sn[s] = br1 * t[s].th1 * i4;
            cn[s] = br2 * t[s].th2 * i4;
            dn[s] = br3 * t[s].th3 * i4;
            Z[s] = bw * t[s].dth4 * i4;
            iden2[s] = 0.0;
        }
    }
    double E[NS], Pi[NS][3];
    {
        const double eokm1 = w(R_TH + TB_EOKM1);
        KG_S_LOOP E[s] = fma(eokm1, X[s], X[s]) + Z[s];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double pa = w(R_TH + TB_SIZE + i * TP_SIZE + TP_A), pp = w(R_TH + TB_SIZE + i * TP_SIZE + TP_P);
            KG_S_LOOP Pi[s][i] = fma(pa, X[s], pp * L[s][i]);
        }
    }
    double I1[NS], I2[NS], Ip[NS], Im[NS];
    if (flags & F_CASE2) {
        const double q1 = w(R_Q0), q2 = w(R_Q1), q3 = w(R_Q2), q4 = w(R_Q3), root = w(R_Q4);
        const double rr41 = q4 - q1, rr43 = q4 - q3;
        const double ho = w(R_Q5), c6 = w(R_Q6), c7 = w(R_Q7), c8 = w(R_Q8), c9 = w(R_Q9), c10 = w(R_Q10);
        const double po0 = w(R_PO0), po1 = w(R_PO1), po2 = w(R_PO2), eo = w(R_EO);
        const double hq = 0.5 * (q1 * q4 + q2 * q3);
        KG_S_LOOP {
            // H = (dr/dtau)/(r - r3) (:600-606)
            const double H = rr41 * (sn[s] * cn[s] * dn[s] * root) * iden2[s] - ho;
            const double Pi_1 = c6 * (Pi[s][0] - po0);
            const double Pi_p = c7 * (Pi[s][1] - po1);
            const double Pi_m = c8 * (Pi[s][2] - po2);
            I1[s] = q3 * (-tau[s]) + rr43 * Pi_1;
            I2[s] = H - hq * (-tau[s]) - root * (E[s] - eo);
            Ip[s] = tau[s] * c9 - Pi_p;
            Im[s] = tau[s] * c10 - Pi_m;
        }
    } else if (flags & F_CASE3) {
        const double rr1 = w(R_Q0), rr2 = w(R_Q1), AA = w(R_Q2), BB = w(R_Q3), sqAB = w(R_Q10);
        const double b2 = BB * rr2, a1_ = AA * rr1;
        const double nsum = b2 + a1_, ndif = b2 - a1_, dsum = BB + AA, ddif = BB - AA;
        KG_S_LOOP rs[s] = fdiv(fma(nsum, cn[s], ndif), fma(dsum, cn[s], ddif));   // :516-519
        // R1 = (Pi - al f1)/(1 - al^2), f1 = (p1/2) ln|(p1 dn + sn)/(p1 dn - sn)| (B62; :666-674)
        double R1a[NS][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double al = w(R_Q4 + i), p1 = w(R_Q7 + i), c0 = w(R_C0 + i);
            KG_S_LOOP {
                const double f1 = p1 * atanh_ratio(p1 * dn[s], sn[s]);
                R1a[s][i] = (Pi[s][i] - al * f1) * c0;
            }
        }
        // R2 (B65; :675-683)
        const double al0 = w(R_Q4), al2 = al0 * al0, iden = w(R_C9), c10 = w(R_C10), n0 = w(R_N0);
        const double po0 = w(R_PO0), po1 = w(R_PO1), po2 = w(R_PO2), eo = w(R_EO);
        const double f0 = w(R_C3), c4 = w(R_C4), c5 = w(R_C5), pr = w(R_C6), c7 = w(R_C7), c8 = w(R_C8);
        const bool lowspin = (flags & F_LOWSPIN) != 0;
        KG_S_LOOP {
            double R2a = (X[s] - (al2 * iden) * (E[s] - fdiv(al0 * sn[s] * dn[s], 1.0 + al0 * cn[s]))) * c10;
            R2a += (2.0 * k - n0) * R1a[s][0] * iden;
            const double dm = lowspin ? 0.0 : (R1a[s][2] - po2);              // :529-534
            const double Pi_1 = f0 * (R1a[s][0] - po0);
            const double Pi_2 = sq(f0) * (R2a - eo);
            const double Pi_p = c4 * (R1a[s][1] - po1);
            const double Pi_m = c5 * dm;
            I1[s] = pr * (-tau[s]) + Pi_1;
            I2[s] = pr * pr * (-tau[s]) + 2.0 * pr * Pi_1 + sqAB * Pi_2;
            Ip[s] = (dsum * (-tau[s]) + Pi_p) * c7;
            Im[s] = (dsum * (-tau[s]) + Pi_m) * c8;
        }
    } else {  // case 4
        const double a2 = w(R_Q0), b1 = w(R_Q1), g0 = w(R_Q2);
        KG_S_LOOP rs[s] = fma(-a2, fdiv(fma(g0, cn[s], -sn[s]), fma(g0, sn[s], cn[s])), -b1);   // :441 with sc = sn/cn
        // S1 = (F + al^2 Pi - al f2)/(1 + al^2), f2 = (p2/2) ln|(1-p2)(1+p2 dn)/((1+p2)(1-p2 dn))| (B92; :638-650)
        double S1a[NS][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double al = w(R_Q2 + i), p2 = w(R_Q5 + i), c3 = w(R_C3 + i), c0 = w(R_C0 + i);
            KG_S_LOOP {
                const double f2 = p2 * (atanh_ratio(1.0, p2 * dn[s]) + c3);
                S1a[s][i] = (X[s] + al * al * Pi[s][i] - al * f2) * c0;
            }
        }
        // S2 (B93; :651-663)
        const double al2 = g0 * g0, i1 = w(R_C0), i2 = w(R_C11);
        const double pref = w(R_Q8), a2g = w(R_C6), c7 = w(R_C7), c8 = w(R_C8), c9 = w(R_C9), c10 = w(R_C10);
        const double po0 = w(R_PO0), po1 = w(R_PO1), po2 = w(R_PO2), eo = w(R_EO);
        const double aa0 = a2g * (1.0 + al2), aa1 = a2g - b1;
        KG_S_LOOP {
            double S2a = (1.0 - k) * X[s] + al2 * E[s] + fdiv(al2 * dn[s] * (g0 * cn[s] - sn[s]), cn[s] + g0 * sn[s]) - al2 * g0;
            S2a = -S2a * (i1 * i2) + S1a[s][0] * (i1 + (1.0 - k) * i2);
            const double Pi_1 = pref * aa0 * (S1a[s][0] - po0);
            const double Pi_2 = pref * (aa0 * aa0) * (S2a - eo);
            const double Pi_p = c7 * (S1a[s][1] - po1);
            const double Pi_m = c8 * (S1a[s][2] - po2);
            I1[s] = aa1 * (-tau[s]) - Pi_1;
            I2[s] = (aa1 * aa1) * (-tau[s]) - 2.0 * aa1 * Pi_1 + Pi_2;
            Ip[s] = c9 * (-tau[s] - Pi_p);
            Im[s] = c10 * (-tau[s] - Pi_m);
        }
    }
    // kgeo/kerr_raytracing_ana.py:626-631
    const double rp = o.rplus, rm = o.rminus, a = o.a, lam = w(S_LAM);
    const double hal = 0.5 * a * lam;
    KG_S_LOOP {
        Iph[s] = o.c_iph * ((rp - hal) * Ip[s] - (rm - hal) * Im[s]);
        const double ItA = o.c_it * ((rp * rp - hal * rp) * Ip[s] - (rm * rm - hal * rm) * Im[s]);
        It[s] = ItA + 4.0 * (-tau[s]) + 2.0 * I1[s] + I2[s];
        Isig[s] = I2[s];
    }
}

}  // namespace kg
