"""R&D (CPU, mpmath): the case-3 / case-4 integrals R1, S1 of GL19a (kgeo/kerr_raytracing_ana.py:638-683) as ONE logarithm.

    R1(al; phi | j) = [Pi(n; phi|j) - al f1] / (1 - al^2),   n = al^2/(al^2-1) > 1,   f1 = (p1/2) ln|(p1 dn + sn)/(p1 dn - sn)|
    S1(al; phi | j) = [F + al^2 Pi(n; phi|j) - al f2] / (1 + al^2),   n = 1 + al^2,   f2 = (p2/2) ln|(1-p2)(1+p2 dn)/((1+p2)(1-p2 dn))|

With Jacobi's theta form  Pi(n; am u|j) = A u + P L,  L = (1/2) ln|th1(v-vb)/th1(v+vb)|  (n > 1: sn^2 b = 1/n, P = -sqrt(n)/sqrt((n-1)(n-j)))
one finds P = -|al| p1 (case 3) and al^2 P = -|al| p2 (case 4), and the product formula
    th1(x+y) th1(x-y) th4(0)^2 = th1(x)^2 th4(y)^2 - th4(x)^2 th1(y)^2
turns (p1 dn)^2 - sn^2 and 1 - (p2 dn)^2 into  const * th1(v+vb) th1(v-vb) / th4(v)^2.  Hence, with N3 = p1 R3 th3(v) + R1 th1(v) and
N4 = th4(v) + p2 R3 th3(v) (R1 = th3(0)/th2(0), R3 = th4(0)/th3(0)):
    L + l = ln|N / th1(v+vb)| + const,      L - l = ln|th1(v-vb) / N| + const       (l = the f1 / f2 logarithm without its prefactor)
i.e. Pi - al f1 = A u - |al| p1 (L + sgn(al) l): ONE logarithm, and for sgn(al) > 0 no zero of its argument on the real axis (the
"removable singularity" of the two-logarithm form is gone).  This script checks both identities to 30 digits."""
import mpmath as mp
mp.mp.dps = 40


def thetas(v, q):
    return [mp.jtheta(i, v, q) for i in (1, 2, 3, 4)]


def check(case, al, j, frac):
    al, j = mp.mpf(al), mp.mpf(j)
    K, Kp = mp.ellipk(j), mp.ellipk(1 - j)
    q = mp.exp(-mp.pi * Kp / K)
    u = frac * K
    v = mp.pi * u / (2 * K)
    sn, cn, dn = [mp.ellipfun(f, u, m=j) for f in ('sn', 'cn', 'dn')]
    t1, t2, t3, t4 = thetas(v, q)
    z1, z2, z3, z4 = thetas(0, q)
    R1c, R3c = z3 / z2, z4 / z3
    n = al * al / (al * al - 1) if case == 3 else 1 + al * al
    b = mp.ellipf(mp.asin(1 / mp.sqrt(n)), j)
    vb = mp.pi * b / (2 * K)
    L = mp.log(abs(mp.jtheta(1, v - vb, q) / mp.jtheta(1, v + vb, q))) / 2
    if case == 3:
        p = mp.sqrt((al * al - 1) / (j + (1 - j) * al * al))
        l = mp.log(abs((p * dn + sn) / (p * dn - sn))) / 2
        N = p * R3c * t3 + R1c * t1
        P = -mp.sqrt(n) / mp.sqrt((n - 1) * (n - j))
        assert abs(P + abs(al) * p) < mp.mpf(10) ** -30
    else:
        p = mp.sqrt((1 + al * al) / (1 - j + al * al))
        l = mp.log(abs((1 + p * dn) / (1 - p * dn))) / 2
        N = t4 + p * R3c * t3
        P = -mp.sqrt(n) / mp.sqrt((n - 1) * (n - j))
        assert abs(al * al * P + abs(al) * p) < mp.mpf(10) ** -30
    plus = L + l - mp.log(abs(N / mp.jtheta(1, v + vb, q)))
    minus = L - l - mp.log(abs(mp.jtheta(1, v - vb, q) / N))
    return plus, minus


for case in (3, 4):
    for al, j in ((1.7, 0.3), (-2.4, 0.8), (5.0, 0.05), (0.6, 0.5) if case == 4 else (1.05, 0.5)):
        vals = [check(case, al, j, f) for f in (0.13, 0.7, 1.3, 2.9)]
        dplus = max(abs(vals[i][0] - vals[0][0]) for i in range(4))
        dminus = max(abs(vals[i][1] + vals[0][0]) for i in range(4))    # the two constants are opposite
        print("case %d al=%5.2f j=%.2f: L+l - ln|N/th1(v+vb)| constant to %s; L-l - ln|th1(v-vb)/N| = -that constant to %s" %
              (case, al, j, mp.nstr(dplus, 3), mp.nstr(dminus, 3)))
