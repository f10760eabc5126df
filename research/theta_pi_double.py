"""Double-precision theta-series evaluation of Pi(n; am(u|m) | m), E(am u|m), sn/cn/dn for per-(n,m) precomputed constants.
Checks accuracy vs mpmath for the four cases."""
import numpy as np, mpmath as mp, random
mp.mp.dps = 40
J = 9

def setup(n, m):
    """per-(n,m) constants (computed in high precision here; the CUDA setup will use fp64 Carlson)"""
    n, m = mp.mpf(n), mp.mpf(m)
    K, Kp, Ec = mp.ellipk(m), mp.ellipk(1 - m), mp.ellipe(m)
    q = mp.exp(-mp.pi * Kp / K)
    c = dict(m=float(m), K=float(K), q=float(q), EoK=float(Ec / K), w=float(mp.pi / (2 * K)))
    m1 = 1 - m
    if 0 < n < m:
        a = mp.ellipf(mp.asin(mp.sqrt(n / m)), m)
        s, cc, d = [mp.ellipfun(f, a, m=m) for f in ('sn', 'cn', 'dn')]
        Za = mp.ellipe(mp.asin(s), m) - (Ec / K) * a
        c.update(kind='A', P=float(s / (cc * d)), Zs=float(Za), sh=float(mp.pi * a / (2 * K)))
    elif n > 1:
        b = mp.ellipf(mp.asin(1 / mp.sqrt(n)), m)
        s, cc, d = [mp.ellipfun(f, b, m=m) for f in ('sn', 'cn', 'dn')]
        Zb = mp.ellipe(mp.asin(s), m) - (Ec / K) * b
        c.update(kind='B', P=float(-s / (cc * d)), Zs=float(Zb + cc * d / s), sh=float(mp.pi * b / (2 * K)))
    else:
        if n < 0:
            # sc^2(b|m') = -n/m
            b = mp.ellipf(mp.atan(mp.sqrt(-n / m)), m1)
        else:
            # dn^2(b|m') = m/n  -> sn^2 = (1 - m/n)/m'
            b = mp.ellipf(mp.asin(mp.sqrt((1 - m / n) / m1)), m1)
        s, cc, d = [mp.ellipfun(f, b, m=m1) for f in ('sn', 'cn', 'dn')]
        Zbp = mp.ellipe(mp.asin(s), m1) - (mp.ellipe(m1) / Kp) * b
        Zd = d * s / cc - Zbp - mp.pi * b / (2 * K * Kp)
        if n < 0:
            c.update(kind='D', P=float(-(s * cc / d)), Zs=float(Zd), sh=float(mp.pi * b / (2 * K)))
        else:
            c.update(kind='C', P=float(-(d / (m1 * s * cc))), Zs=float(Zd - m * s / (cc * d)), sh=float(mp.pi * b / (2 * K)))
    return c

def evaluate(c, u):
    """returns Pi, E, sn, cn, dn at u (float64 arithmetic)"""
    q, w = c['q'], c['w']
    v = w * u
    cv, sv = np.cos(v), np.sin(v)
    c1, s1 = cv * cv - sv * sv, 2 * sv * cv          # cos 2v, sin 2v
    cj, sj = [1.0, c1], [0.0, s1]
    for j in range(2, J + 1):
        cj.append(2 * c1 * cj[-1] - cj[-2]); sj.append(2 * c1 * sj[-1] - sj[-2])
    qj2 = [q ** (j * j) for j in range(J + 1)]
    th4 = 1 + 2 * sum((-1) ** j * qj2[j] * cj[j] for j in range(1, J + 1))
    th3 = 1 + 2 * sum(qj2[j] * cj[j] for j in range(1, J + 1))
    dth4 = -4 * sum((-1) ** j * j * qj2[j] * sj[j] for j in range(1, J + 1))
    # odd harmonics sin((2j+1)v), cos((2j+1)v)
    so = [sj[j] * cv + cj[j] * sv for j in range(J)]
    co = [cj[j] * cv - sj[j] * sv for j in range(J)]
    qh = [q ** ((j + 0.5) ** 2) for j in range(J)]
    th1 = 2 * sum((-1) ** j * qh[j] * so[j] for j in range(J))
    th2 = 2 * sum(qh[j] * co[j] for j in range(J))
    t2_0 = 2 * sum(qh[j] for j in range(J)); t3_0 = 1 + 2 * sum(qj2[j] for j in range(1, J + 1)); t4_0 = 1 + 2 * sum((-1) ** j * qj2[j] for j in range(1, J + 1))
    sn = (t3_0 / t2_0) * th1 / th4; cn = (t4_0 / t2_0) * th2 / th4; dn = (t4_0 / t3_0) * th3 / th4
    Z = w * dth4 / th4
    E = c['EoK'] * u + Z
    sh, kind = c['sh'], c['kind']
    if kind == 'A':      # real shift, theta4:  L = atanh(O/E)
        Ee = 1 + 2 * sum((-1) ** j * qj2[j] * np.cos(2 * j * sh) * cj[j] for j in range(1, J + 1))
        Oo = 2 * sum((-1) ** j * qj2[j] * np.sin(2 * j * sh) * sj[j] for j in range(1, J + 1))
        L = 0.5 * np.log((Ee + Oo) / (Ee - Oo))
    elif kind == 'B':    # real shift, theta1: theta1(v -+ sh) = 2 sum (-1)^j qh (so cos((2j+1)sh) -+ co sin((2j+1)sh))
        Ee = 2 * sum((-1) ** j * qh[j] * np.cos((2 * j + 1) * sh) * so[j] for j in range(J))
        Oo = 2 * sum((-1) ** j * qh[j] * np.sin((2 * j + 1) * sh) * co[j] for j in range(J))
        L = 0.5 * np.log(abs((Ee - Oo) / (Ee + Oo)))
    elif kind == 'D':    # imaginary shift, theta4
        Re = 1 + 2 * sum((-1) ** j * qj2[j] * np.cosh(2 * j * sh) * cj[j] for j in range(1, J + 1))
        Im = 2 * sum((-1) ** j * qj2[j] * np.sinh(2 * j * sh) * sj[j] for j in range(1, J + 1))
        L = np.arctan2(Im, Re)
    else:                # 'C': imaginary shift, theta3
        Re = 1 + 2 * sum(qj2[j] * np.cosh(2 * j * sh) * cj[j] for j in range(1, J + 1))
        Im = 2 * sum(qj2[j] * np.sinh(2 * j * sh) * sj[j] for j in range(1, J + 1))
        L = np.arctan2(Im, Re)
    Pi = u + c['P'] * (u * c['Zs'] + L)
    return Pi, E, sn, cn, dn

random.seed(11)
for case in ('A', 'B', 'C', 'D'):
    worst = dict(Pi=0, E=0, sn=0)
    for _ in range(40):
        m = random.uniform(0.02, 0.98)
        n = {'A': m * random.uniform(0.03, 0.97), 'B': 1 + random.uniform(0.03, 8), 'C': m + (1 - m) * random.uniform(0.03, 0.97),
             'D': -random.uniform(0.03, 8)}[case]
        c = setup(n, m)
        assert c['kind'] == case
        K = c['K']
        for frac in (0.13, 0.8, 1.9, 3.3, -2.4):
            u = frac * K
            Pi, E, sn, cn, dn = evaluate(c, u)
            mm, nn, uu = mp.mpf(m), mp.mpf(n), mp.mpf(u)
            sn_t = mp.ellipfun('sn', uu, m=mm)
            if case == 'B':
                if abs(1 - nn * sn_t ** 2) < 1e-2: continue
                # truth through the high-precision theta formula (verified separately)
                Kt, Kp = mp.ellipk(mm), mp.ellipk(1 - mm); qq = mp.exp(-mp.pi * Kp / Kt)
                b = mp.ellipf(mp.asin(1 / mp.sqrt(nn)), mm); s, cc, d = [mp.ellipfun(f, b, m=mm) for f in ('sn', 'cn', 'dn')]
                vb, v = mp.pi * b / (2 * Kt), mp.pi * uu / (2 * Kt)
                Zb = (mp.pi / (2 * Kt)) * mp.jtheta(4, vb, qq, 1) / mp.jtheta(4, vb, qq)
                want = uu - (s / (cc * d)) * (uu * (Zb + cc * d / s) + mp.log(abs(mp.jtheta(1, v - vb, qq) / mp.jtheta(1, v + vb, qq))) / 2)
            else:
                want = mp.quad(lambda t: 1 / (1 - nn * mp.ellipfun('sn', t, m=mm) ** 2), mp.linspace(0, uu, 12))
            am = mp.quad(lambda t: mp.ellipfun('dn', t, m=mm), mp.linspace(0, uu, 12))
            E_t = mp.quad(lambda t: mp.ellipfun('dn', t, m=mm) ** 2, mp.linspace(0, uu, 12))
            ePi = abs(Pi - want) / max(abs(want), abs(uu)); eE = abs(E - E_t) / abs(uu); es = abs(sn - sn_t)
            if ePi > worst['Pi']: worst['Pi'] = ePi; wp = (m, n, frac, c['q'], c['P'])
            worst['E'] = max(worst['E'], eE); worst['sn'] = max(worst['sn'], es)
    print(case, {k: float(v) for k, v in worst.items()}, 'worst at m,n,frac,q,P', wp)
