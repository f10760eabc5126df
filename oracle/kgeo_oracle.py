"""CPU oracle for the kgeo analytic Kerr ray-tracing hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the algorithm of the reference
(achael/kgeo, mounted read-only at /root/reference during development).  It is the
checker the CUDA path is compared with; it is NOT part of the product.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  `kgeo_b200/` never does.

Third-party arithmetic: like the reference, the oracle calls `scipy.special`
(`elliprf/elliprj/elliprc/ellipkinc/ellipeinc/ellipj/ellipk`); the reference does not
pin SciPy (setup.py:24-31); the build container and the GPU box both carry
scipy 1.18.1 / numpy 2.3.5.

Pinning: `tests/test_oracle_vs_golden.py` checks every function below against
fixtures produced by running the UNMODIFIED reference in the build container
(`tests/golden/make_golden.py`, fixtures `tests/golden/*.npz`) and against the
reference's own known-answer tests (ipole rays, velocity / B-field / emissivity
tables, EVPA CSVs copied as data fixtures).

Every function cites the reference file:line it restates (paths relative to
/root/reference/).
"""
import numpy as np
import scipy.special as sp

MINSPIN = 1.0e-4   # kgeo/kerr_raytracing_utils.py:9
EP = 1.0e-12       # kgeo/kerr_raytracing_utils.py:12
HALFPI = 0.5 * np.pi


def _arr(x):
    return np.atleast_1d(np.asarray(x, dtype=float))


# --------------------------------------------------------------------------------------
# Legendre Pi(n; phi | m) through Carlson RF/RJ/RC   (kgeo/scipy_ellip_binding.py:87-237)
# --------------------------------------------------------------------------------------
def _pi_principal(n, phi, m):
    """Carlson evaluation valid for |phi| <= pi/2 (scipy_ellip_binding.py:194-235)."""
    s = np.sin(phi)
    c = np.cos(phi)
    x = c ** 2
    y = 1.0 - m * s ** 2          # operation order as in the reference (:116-119): m*(s^2)
    z = np.ones_like(x)
    rho = 1.0 - n * s ** 2
    rf = sp.elliprf(x, y, z)
    rj = np.zeros_like(x)
    pos = rho > 0
    if pos.any():
        rj[pos] = sp.elliprj(x[pos], y[pos], z[pos], rho[pos])
    neg = ~pos
    if neg.any():
        # DLMF 19.20.14 (Cauchy principal value for negative 4th argument), :207-230
        xn, yn, zn = x[neg], y[neg], z[neg]
        q = -rho[neg]
        p = (zn * (xn + yn + q) - xn * yn) / (zn + q)
        rj0 = sp.elliprj(xn, yn, zn, p)
        rc = sp.elliprc(xn * yn + p * q, p * q)
        w = 3.0 * np.sqrt((xn * yn * zn) / (xn * yn + p * q))
        t = ((p - zn) * rj0 - 3.0 * rf[neg] + w * rc)
        rj[neg] = t / (q + zn)
    return s * rf + (n / 3.0) * (s ** 3) * rj


def legendre_pi(n, phi, m):
    """Pi(n; phi | m), A&S / Mathematica convention, any real phi.

    Restates ellip_pi_arr (kgeo/scipy_ellip_binding.py:87-237): the seven special-value
    masks are applied in the same order (:125-162), then the two periodicity
    reductions (:164-192), each of which re-enters the full function.
    `n` and `m` broadcast against `phi` (the reference tiles them with np.outer, :103-106).
    """
    phi = np.asarray(phi, dtype=float)
    shape = phi.shape
    n_b = np.broadcast_to(np.asarray(n, dtype=float), shape).ravel().copy()
    m_b = np.broadcast_to(np.asarray(m, dtype=float), shape).ravel().copy()
    out = _legendre_pi_flat(n_b, phi.ravel().copy(), m_b)
    return out.reshape(shape)


def _legendre_pi_flat(n, phi, m):
    out = np.zeros_like(phi)
    todo = np.ones(phi.shape, dtype=bool)
    s = np.sin(phi)
    y = 1.0 - m * s ** 2
    rho = 1.0 - n * s ** 2
    aphi = np.abs(phi)

    def settle(mask, value):
        mask = mask & todo
        if mask.any():
            out[mask] = value[mask] if isinstance(value, np.ndarray) else value
        todo[mask] = False

    settle(aphi <= 1.0e-10, phi)                    # :126
    settle(m < -1.0e14, 0.0)                        # :132
    settle(y < 0, np.nan)                           # :138
    settle((aphi > HALFPI) & (m >= 1), np.nan)      # :143
    settle((aphi > HALFPI) & (n == 1), np.inf)      # :149
    settle(rho == 0, np.inf)                        # :154
    settle(np.isinf(n), 0.0)                        # :159

    far = (aphi > np.pi) & todo                     # :165-178
    if far.any():
        sg = np.sign(phi[far])
        k = sg * np.floor(aphi[far] / np.pi)
        rest = phi[far] - k * np.pi
        comp = _legendre_pi_flat(n[far], np.full(rest.shape, HALFPI), m[far])
        out[far] = 2.0 * k * comp + _legendre_pi_flat(n[far], rest, m[far])
        todo[far] = False

    mid = (aphi > HALFPI) & (aphi <= np.pi) & todo  # :180-192
    if mid.any():
        sg = np.sign(phi[mid])
        refl = -phi[mid] + sg * np.pi
        comp = _legendre_pi_flat(n[mid], np.full(refl.shape, HALFPI), m[mid])
        out[mid] = sg * 2.0 * comp - _legendre_pi_flat(n[mid], refl, m[mid])
        todo[mid] = False

    if todo.any():
        out[todo] = _pi_principal(n[todo], phi[todo], m[todo])
    return out


# --------------------------------------------------------------------------------------
# per-ray constants                                  (kgeo/kerr_raytracing_utils.py)
# --------------------------------------------------------------------------------------
def conserved(a, th_o, alpha, beta):
    """lam, eta from screen coordinates (kgeo/kerr_raytracing_ana.py:76-77)."""
    lam = -alpha * np.sin(th_o)
    eta = (alpha ** 2 - a ** 2) * np.cos(th_o) ** 2 + beta ** 2
    return lam, eta


def sign_pos0(x):
    """sign with sign(0)=+1 (kerr_raytracing_utils.py:857-864)."""
    return np.where(np.asarray(x) >= 0, 1.0, -1.0)


def angular_consts(a, th_o, lam, eta):
    """u_plus, u_minus, u_plus/u_minus, a^2 u_minus (kerr_raytracing_utils.py:276-328)."""
    lam = _arr(lam).copy()
    eta = _arr(eta).copy()
    c2 = np.cos(th_o) ** 2
    if abs(a) < MINSPIN:                                  # :290-292
        up = eta / (eta + lam * lam)
        um = up.copy()
    else:                                                 # :295-297
        dth = 0.5 * (1.0 - (eta + lam * lam) / (a * a))
        root = np.sqrt(dth ** 2 + eta / (a * a))
        up = dth + root
        um = dth - root
    up = np.where((c2 - up) > 0, c2, up)                  # :300-301
    up = np.where(up ** 2 > 1, 1.0, up)                   # :304-305
    um = np.where((um == 0) & (eta < 0), EP, um)          # :311
    um = np.where((um == 0) & (eta >= 0), -EP, um)        # :312
    um = np.where((eta < 0) & ((c2 - um) < 0), c2, um)    # :315-318
    if abs(a) < MINSPIN:                                  # :321-326
        ratio = 0.0 * um
        a2um = -(eta + lam ** 2)
    else:
        ratio = up / um
        a2um = a ** 2 * um
    return up, um, ratio, a2um


def angular_turning(a, th_o, lam, eta):
    """u_plus, u_minus, theta_plus, theta_minus, thclass (kerr_raytracing_utils.py:242-274)."""
    lam = _arr(lam)
    eta = _arr(eta)
    up, um, _, _ = angular_consts(a, th_o, lam, eta)
    th_plus = np.arccos(-np.sqrt(up))                     # :257 (GL19a theta_4)
    th_minus = np.arccos(np.sqrt(up))                     # :258 (GL19a theta_1)
    th_plus = np.where(eta < 0, np.arccos(np.sqrt(np.where(eta < 0, um, 0.0))), th_plus)   # :262 vortical: theta_2
    thclass = np.zeros(lam.shape, dtype=int)              # :265-272
    thclass[eta > 0] = 1
    thclass[eta < 0] = 2
    thclass[eta == 0] = 3
    return up, um, th_plus, th_minus, thclass


def _cbrt_realmost(x):
    """Complex cube root with the largest real part (kerr_raytracing_utils.py:828-855)."""
    x = np.asarray(x, dtype=complex)
    mag = np.cbrt(np.abs(x))
    ang = np.angle(x)
    out = np.empty(x.shape, dtype=complex)
    isreal = (x.imag == 0)
    out[isreal] = (np.sign(x) * mag)[isreal]
    if (~isreal).any():
        cands = [mag * np.exp(1j * ang / 3.0 + 1j * sh) for sh in (0.0, 2.0 * np.pi / 3.0, -2.0 * np.pi / 3.0)]
        for i, ci in enumerate(cands):
            others = [cands[j] for j in range(3) if j != i]
            pick = (~isreal) & (ci.real > others[0].real) & (ci.real > others[1].real)
            out[pick] = ci[pick]
    return out


def quartic_roots(a, lam, eta):
    """Roots r1..r4 of the radial potential (complex128) and region label 1..4
    (kerr_raytracing_utils.py:330-380, GL19a eq. 95)."""
    lam = _arr(lam)
    eta = _arr(eta)
    rplus = 1.0 + np.sqrt(1.0 - a ** 2)
    A = a ** 2 - eta - lam ** 2
    B = 2.0 * (eta + (lam - a) ** 2)
    C = -(a ** 2) * eta
    P = -(A ** 2) / 12.0 - C
    Q = (-A / 3.0) * ((A ** 2) / 36.0 - C) - (B ** 2) / 8.0
    disc = (4.0 * (P ** 3) + 27.0 * (Q ** 2)).astype(complex)
    half = np.sqrt(disc / 108.0)
    wp = _cbrt_realmost(-0.5 * Q + half)
    wm = _cbrt_realmost(-0.5 * Q - half)
    xi0 = wp + wm - A / 3.0
    z = np.sqrt(0.5 * xi0)
    d12 = (-0.5 * A - z ** 2 + 0.25 * B / z).astype(complex)
    d34 = (-0.5 * A - z ** 2 - 0.25 * B / z).astype(complex)
    r1 = -z - np.sqrt(d12)
    r2 = -z + np.sqrt(d12)
    r3 = z - np.sqrt(d34)
    r4 = z + np.sqrt(d34)
    region = np.zeros(r1.shape, dtype=int)
    region[r2.imag != 0] = 4
    region[(r2.imag == 0) & (r4.imag != 0)] = 3
    region[(r2.imag == 0) & (r4.imag == 0) & (r4.real < rplus)] = 2
    region[(r2.imag == 0) & (r4.imag == 0) & (r4.real > rplus)] = 1
    return r1, r2, r3, r4, region


def _case4_params(a, r_o, r2, r4):
    """Shared parameters of radial case 4 (kerr_raytracing_ana.py:410-434; utils :418-438)."""
    rplus = 1.0 + np.sqrt(1.0 - a ** 2)
    rminus = 1.0 - np.sqrt(1.0 - a ** 2)
    a1, b1, a2, b2 = r4.imag, r4.real, r2.imag, r2.real
    CC = np.sqrt((a1 - a2) ** 2 + (b1 - b2) ** 2)
    DD = np.sqrt((a1 + a2) ** 2 + (b1 - b2) ** 2)
    k4 = (4.0 * CC * DD) / ((CC + DD) ** 2)
    x4rp = (rplus + b1) / a2
    x4rm = (rminus + b1) / a2
    x4ro = np.full(a2.shape, np.inf) if np.isinf(r_o) else (r_o + b1) / a2
    g0 = np.sqrt((4.0 * a2 ** 2 - (CC - DD) ** 2) / ((CC + DD) ** 2 - 4.0 * a2 ** 2))
    return dict(a1=a1, b1=b1, a2=a2, b2=b2, CC=CC, DD=DD, k4=k4, x4rp=x4rp, x4rm=x4rm, x4ro=x4ro, g0=g0)


def _case3_params(a, r_o, r1, r2, r4):
    """Shared parameters of radial case 3 (kerr_raytracing_ana.py:479-505; utils :451-473)."""
    rplus = 1.0 + np.sqrt(1.0 - a ** 2)
    rminus = 1.0 - np.sqrt(1.0 - a ** 2)
    a1, b1 = r4.imag, r4.real
    rr1, rr2 = r1.real, r2.real
    AA = np.sqrt(a1 ** 2 + (b1 - rr2) ** 2)
    BB = np.sqrt(a1 ** 2 + (b1 - rr1) ** 2)
    k3 = ((AA + BB) ** 2 - (rr2 - rr1) ** 2) / (4.0 * AA * BB)
    rrp1, rrp2 = rplus - rr1, rplus - rr2
    rrm1, rrm2 = rminus - rr1, rminus - rr2
    x3rp = (AA * rrp1 - BB * rrp2) / (AA * rrp1 + BB * rrp2)
    x3rm = (AA * rrm1 - BB * rrm2) / (AA * rrm1 + BB * rrm2)
    if np.isinf(r_o):
        x3ro = (AA - BB) / (AA + BB)
    else:
        x3ro = (AA * (r_o - rr1) - BB * (r_o - rr2)) / (AA * (r_o - rr1) + BB * (r_o - rr2))
    return dict(rr1=rr1, rr2=rr2, rr21=rr2 - rr1, rrp1=rrp1, rrp2=rrp2, rrm1=rrm1, rrm2=rrm2,
                AA=AA, BB=BB, k3=k3, x3rp=x3rp, x3rm=x3rm, x3ro=x3ro)


def _case2_x2ro(r_o, rr31, rr41, rr3, rr4):
    if np.isinf(r_o):
        return np.sqrt(rr31 / rr41)
    return np.sqrt((rr31 * (r_o - rr4)) / (rr41 * (r_o - rr3)))


def mino_total(a, r_o, eta, r1, r2, r3, r4):
    """Total Mino time along each ray (kerr_raytracing_utils.py:382-548)."""
    rplus = 1.0 + np.sqrt(1.0 - a ** 2)
    out = np.zeros(r1.shape)
    m4 = (r1.imag != 0)                                           # :412
    if m4.any():
        p = _case4_params(a, r_o, r2[m4], r4[m4])
        pref = 2.0 / (p['CC'] + p['DD'])
        hor = pref * sp.ellipkinc(np.arctan(p['x4rp']) + np.arctan(p['g0']), p['k4'])
        obs = pref * sp.ellipkinc(np.arctan(p['x4ro']) + np.arctan(p['g0']), p['k4'])
        out[m4] = obs - hor                                       # :441
    m3 = (r3.imag != 0) & ~m4                                     # :445
    if m3.any():
        p = _case3_params(a, r_o, r1[m3], r2[m3], r4[m3])
        pref = 1.0 / np.sqrt(p['AA'] * p['BB'])
        hor = pref * sp.ellipkinc(np.arccos(p['x3rp']), p['k3'])
        obs = pref * sp.ellipkinc(np.arccos(p['x3ro']), p['k3'])
        out[m3] = obs - hor                                       # :481
    with np.errstate(invalid='ignore'):
        m2 = (r3.imag == 0) & (r3.real < rplus) & ~m3             # :485 (complex '<' compares real parts)
        m1 = (r3.imag == 0) & (r3.real > rplus) & ~m2             # :518
    for mask, turning in ((m2, False), (m1, True)):
        if not mask.any():
            continue
        q1, q2, q3, q4 = r1[mask].real, r2[mask].real, r3[mask].real, r4[mask].real
        rr31, rr32, rr41, rr42 = q3 - q1, q3 - q2, q4 - q1, q4 - q2
        k2 = (rr32 * rr41) / (rr31 * rr42)
        pref = 2.0 / np.sqrt(rr31 * rr42)
        obs = pref * sp.ellipkinc(np.arcsin(_case2_x2ro(r_o, rr31, rr41, q3, q4)), k2)
        if turning:
            out[mask] = 2.0 * obs                                 # :545
        else:
            x2rp = np.sqrt((rr31 * (rplus - q4)) / (rr41 * (rplus - q3)))
            out[mask] = obs - pref * sp.ellipkinc(np.arcsin(x2rp), k2)   # :514
    return out


# --------------------------------------------------------------------------------------
# theta(tau), G_phi, G_t                          (kgeo/kerr_raytracing_ana.py:195-365)
# --------------------------------------------------------------------------------------
def _jacobi_amplitude_negative_parameter(arg, k):
    """sin(am) and am of argument `arg` for NEGATIVE parameter k via the reciprocal-modulus
    identities the reference uses (kerr_raytracing_ana.py:262-277 / :337-352)."""
    sin_am = np.zeros(arg.shape)
    am = np.zeros(arg.shape)
    tiny = np.abs(arg) < EP
    sin_am[tiny] = arg[tiny]
    am[tiny] = arg[tiny]
    big = ~tiny
    if big.any():
        m = np.broadcast_to(k / (k - 1.0), arg.shape)[big]
        scaled = arg[big] / np.sqrt(1.0 - m)
        sn, cn, dn, _ = sp.ellipj(scaled, m)
        sin_am[big] = np.sqrt(1.0 - m) * sn / dn
        am[big] = 0.5 * np.pi - sp.ellipj(sp.ellipk(m) - scaled, m)[3]
    return sin_am, am


def theta_motion(a, th_o, s_o, lam, eta, up, um, tau, full=True):
    """theta_s, G_phi, G_t at Mino times tau[ntau, npix] (kerr_raytracing_ana.py:195-365)."""
    th_s = np.zeros(tau.shape)
    G_ph = np.zeros(tau.shape)
    G_t = np.zeros(tau.shape)
    cth = np.cos(th_o)

    if np.any(eta > 0.0):                                   # ordinary motion, :217-296
        sel = eta >= 0.0
        p, q, s = up[sel], um[sel], s_o[sel]
        if abs(a) < MINSPIN:
            k = np.zeros(eta.shape)[sel]
            a2um = -(eta + lam ** 2)[sel]
        else:
            k = p / q
            a2um = a ** 2 * q
        pref = 1.0 / np.sqrt(-a2um)
        amp_o = np.arcsin(cth / np.sqrt(p))
        F_o = sp.ellipkinc(amp_o, k)
        Gth_o = -pref * F_o
        if full:
            Gph_o = -pref * legendre_pi(p, amp_o, k)
            Gt_o = np.zeros(k.shape)
            z0 = (k == 0)
            if z0.any():
                Gt_o[z0] = (2 * p * pref * 0.125 * (-2 * amp_o + np.sin(2 * amp_o)))[z0]
            if (~z0).any():
                Gt_o[~z0] = (2 * p * pref * (sp.ellipeinc(amp_o, k) - F_o))[~z0] / (2 * k[~z0])
        arg = (np.sqrt(-a2um) * (-tau[:, sel] + s * Gth_o)).astype(float)
        sinPhi, Phi = _jacobi_amplitude_negative_parameter(arg, k)
        th_s[:, sel] = np.arccos(-s * np.sqrt(p) * sinPhi)
        if full:
            G_ph[:, sel] = pref * legendre_pi(p, Phi, k) - s * Gph_o
            z0 = (k == 0)
            Gt = np.zeros(Phi.shape)
            if z0.any():
                Gt[:, z0] = (-2 * p * pref * 0.125 * (-2 * Phi + np.sin(2 * Phi)))[:, z0]
            if (~z0).any():
                Gt[:, ~z0] = -(2 * p * pref * (sp.ellipeinc(Phi, k) - sp.ellipkinc(Phi, k)))[:, ~z0] / (2 * k[~z0])
            G_t[:, sel] = Gt - s * Gt_o

    if np.any(eta <= 0.0):                                  # vortical motion, :300-363
        if abs(a) < MINSPIN:
            raise Exception("below MINSPIN but there are eta<0 points in th_integrate!")
        sel = eta < 0.0
        p, q, s = up[sel], um[sel], s_o[sel]
        a2um = a ** 2 * q
        h = 1.0 if th_o <= np.pi / 2 else -1.0
        prefA = 1.0 / np.sqrt(a2um)
        prefB = prefA / (1.0 - q)
        prefC = np.sqrt(q / a ** 2)
        nu_o = np.arcsin(np.sqrt((cth ** 2 - q) / (p - q)))
        k = 1.0 - p / q
        nchar = (p - q) / (1.0 - q)
        Gth_o = -h * prefA * sp.ellipkinc(nu_o, k)
        if full:
            Gph_o = -h * prefB * legendre_pi(nchar, nu_o, k)
            Gt_o = -h * prefC * sp.ellipeinc(nu_o, k)
        arg = (np.sqrt(a2um) * (-tau[:, sel] + s * Gth_o)).astype(float)
        sinNu, Nu = _jacobi_amplitude_negative_parameter(arg, k)
        th_s[:, sel] = np.arccos(h * np.sqrt(q + (p - q) * sinNu ** 2))
        if full:
            G_ph[:, sel] = prefB * legendre_pi(nchar, Nu, k) - s * Gph_o
            G_t[:, sel] = prefC * sp.ellipeinc(Nu, k) - s * Gt_o
    return th_s, G_ph, G_t


# --------------------------------------------------------------------------------------
# r(tau), I_phi, I_t, I_sigma                     (kgeo/kerr_raytracing_ana.py:368-683)
# --------------------------------------------------------------------------------------
def _S1_S2(al, phi, j, want_s2=True):
    """GL19a B92/B93 (kerr_raytracing_ana.py:638-663)."""
    al2 = al * al
    al2p1 = 1.0 + al2
    al2j = 1.0 - j + al2
    s2phi = np.sqrt(1.0 - j * (np.sin(phi)) ** 2)
    p2 = np.sqrt(al2p1 / al2j)
    f2 = 0.5 * p2 * np.log(np.abs(((1 - p2) * (1 + p2 * s2phi)) / ((1 + p2) * (1 - p2 * s2phi))))
    F = sp.ellipkinc(phi, j)
    S1 = (F + al2 * legendre_pi(al2p1, phi, j) - al * f2) / al2p1
    if not want_s2:
        return S1, np.nan
    E = sp.ellipeinc(phi, j)
    tanphi = np.tan(phi)
    S2a = (1 - j) * F + al2 * E + al2 * s2phi * (al - tanphi) / (1 + al * tanphi) - al2 * al
    S2a = -S2a / (al2p1 * al2j)
    S2b = S1 * (1.0 / al2p1 + (1 - j) / al2j)
    return S1, S2a + S2b


def _R1_R2(al, phi, j, want_r2=True):
    """GL19a B62/B65 (kerr_raytracing_ana.py:666-683)."""
    al2 = al ** 2
    s2phi = np.sqrt(1 - j * np.sin(phi) ** 2)
    p1 = np.sqrt((al2 - 1) / (j + (1 - j) * al2))
    f1 = 0.5 * p1 * np.log(np.abs((p1 * s2phi + np.sin(phi)) / (p1 * s2phi - np.sin(phi))))
    nn = al2 / (al2 - 1)
    R1 = (legendre_pi(nn, phi, j) - al * f1) / (1 - al2)
    if not want_r2:
        return R1, np.nan
    F = sp.ellipkinc(phi, j)
    E = sp.ellipeinc(phi, j)
    R2 = (F - (al2 / (j + (1 - j) * al2)) * (E - al * np.sin(phi) * s2phi / (1 + al * np.cos(phi)))) / (al2 - 1)
    R2 = R2 + (2 * j - nn) * R1 / (j + (1 - j) * al2)
    return R1, R2


def radial_motion(a, r_o, lam, eta, r1, r2, r3, r4, tau, full=True):
    """r_s, I_phi, I_t, I_sigma at Mino times tau[ntau, npix] (kerr_raytracing_ana.py:368-633)."""
    rplus = 1.0 + np.sqrt(1.0 - a ** 2)
    rminus = 1.0 - np.sqrt(1.0 - a ** 2)
    r_s = np.zeros(tau.shape)
    I_p = np.zeros(tau.shape)
    I_m = np.zeros(tau.shape)
    I_1 = np.zeros(tau.shape)
    I_2 = np.zeros(tau.shape)

    m4 = (r1.imag != 0)                                             # :405
    if m4.any():
        t = tau[:, m4]
        p = _case4_params(a, r_o, r2[m4], r4[m4])
        a2, b1, g0, k4 = p['a2'], p['b1'], p['g0'], p['k4']
        x4rp, x4rm = p['x4rp'], p['x4rm']
        gp = (g0 * x4rp - 1.0) / (g0 + x4rp)
        gm = (g0 * x4rm - 1.0) / (g0 + x4rm)
        aux = np.arctan(p['x4ro']) + np.arctan(g0)
        pref = 2.0 / (p['CC'] + p['DD'])
        Ir_o = pref * sp.ellipkinc(aux, k4)
        X4 = (1.0 / pref) * (-t + Ir_o)
        sn, cn, dn, am = sp.ellipj(X4, k4)
        sc = sn / cn
        r_s[:, m4] = -a2 * (g0 - sc) / (1 + g0 * sc) - b1           # :441
        if full:
            S1a0, S2a0 = _S1_S2(g0, am, k4)
            S1b0, S2b0 = _S1_S2(g0, aux, k4)
            S1ap, _ = _S1_S2(gp, am, k4, False)
            S1bp, _ = _S1_S2(gp, aux, k4, False)
            S1am, _ = _S1_S2(gm, am, k4, False)
            S1bm, _ = _S1_S2(gm, aux, k4, False)
            aa0 = (a2 / g0) * (1 + g0 ** 2)
            Pi_1 = pref * aa0 * (S1a0 - S1b0)
            Pi_2 = pref * (aa0 ** 2) * (S2a0 - S2b0)
            Pi_p = pref * (1 + g0 ** 2) / (g0 ** 2 + g0 * x4rp) * (S1ap - S1bp)
            Pi_m = pref * (1 + g0 ** 2) / (g0 ** 2 + g0 * x4rm) * (S1am - S1bm)
            aa1 = (a2 / g0 - b1)
            I_1[:, m4] = aa1 * (-t) - Pi_1
            I_2[:, m4] = (aa1 ** 2) * (-t) - 2 * aa1 * Pi_1 + Pi_2
            I_p[:, m4] = g0 / (a2 - a2 * g0 * x4rp) * (-t - Pi_p)
            I_m[:, m4] = g0 / (a2 - a2 * g0 * x4rm) * (-t - Pi_m)

    m3 = (r3.imag != 0) & ~m4                                       # :474
    if m3.any():
        t = tau[:, m3]
        p = _case3_params(a, r_o, r1[m3], r2[m3], r4[m3])
        AA, BB, k3 = p['AA'], p['BB'], p['k3']
        rr1, rr2, rr21 = p['rr1'], p['rr2'], p['rr21']
        alp = -1.0 / p['x3rp']
        alm = -1.0 / p['x3rm']
        al0 = (BB + AA) / (BB - AA)
        aux = np.arccos(p['x3ro'])
        pref = 1.0 / np.sqrt(AA * BB)
        Ir_o = pref * sp.ellipkinc(aux, k3)
        X3 = (1.0 / pref) * (-t + Ir_o)
        sn, cn, dn, am = sp.ellipj(X3, k3)
        r_s[:, m3] = ((BB * rr2 - AA * rr1) + (BB * rr2 + AA * rr1) * cn) / ((BB - AA) + (BB + AA) * cn)   # :516-519
        if full:
            R1a0, R2a0 = _R1_R2(al0, am, k3)
            R1b0, R2b0 = _R1_R2(al0, aux, k3)
            R1ap, _ = _R1_R2(alp, am, k3, False)
            R1bp, _ = _R1_R2(alp, aux, k3, False)
            if abs(a) > MINSPIN:
                R1am, _ = _R1_R2(alm, am, k3, False)
                R1bm, _ = _R1_R2(alm, aux, k3, False)
            else:
                R1am = np.zeros(R1ap.shape)
                R1bm = np.zeros(R1ap.shape)
            fac = 2 * rr21 * np.sqrt(AA * BB)
            Pi_1 = (fac / (BB ** 2 - AA ** 2)) * (R1a0 - R1b0)
            Pi_2 = (fac / (BB ** 2 - AA ** 2)) ** 2 * (R2a0 - R2b0)
            Pi_p = (fac / (BB * p['rrp2'] - AA * p['rrp1'])) * (R1ap - R1bp)
            Pi_m = (fac / (BB * p['rrm2'] - AA * p['rrm1'])) * (R1am - R1bm)
            pr = (BB * rr2 + AA * rr1) / (BB + AA)
            I_1[:, m3] = pr * (-t) + Pi_1
            I_2[:, m3] = pr ** 2 * (-t) + 2 * pr * Pi_1 + np.sqrt(AA * BB) * Pi_2
            I_p[:, m3] = -((BB + AA) * (-t) + Pi_p) / (BB * p['rrp2'] + AA * p['rrp1'])
            I_m[:, m3] = -((BB + AA) * (-t) + Pi_m) / (BB * p['rrm2'] + AA * p['rrm1'])

    m2 = (r3.imag == 0.0) & ~m3                                     # :555
    if m2.any():
        t = tau[:, m2]
        q1, q2, q3, q4 = r1[m2].real, r2[m2].real, r3[m2].real, r4[m2].real
        rr31, rr32, rr41, rr42, rr43 = q3 - q1, q3 - q2, q4 - q1, q4 - q2, q4 - q3
        rrp3, rrm3, rrp4, rrm4 = rplus - q3, rminus - q3, rplus - q4, rminus - q4
        k2 = (rr32 * rr41) / (rr31 * rr42)
        aux = np.arcsin(_case2_x2ro(r_o, rr31, rr41, q3, q4))
        root = np.sqrt(rr31 * rr42)
        Ir_o = 2 * sp.ellipkinc(aux, k2) / root
        X2 = 0.5 * root * (-t + Ir_o)
        sn, cn, dn, am = sp.ellipj(X2, k2)
        rs = (q4 * rr31 - q3 * rr41 * sn ** 2) / (rr31 - rr41 * sn ** 2)    # :592-594
        r_s[:, m2] = rs
        if full:
            dsn2 = 2 * sn * cn * dn * (-0.5 * root)
            drs = rr31 * rr43 * rr41 * dsn2 / ((rr31 - rr41 * sn ** 2) ** 2)
            drs = -drs
            Rpot_o = (r_o - q1) * (r_o - q2) * (r_o - q3) * (r_o - q4)
            H = drs / (rs - q3) - np.sqrt(Rpot_o) / (r_o - q3)
            E = root * (sp.ellipeinc(am, k2) - sp.ellipeinc(aux, k2))
            n1 = rr41 / rr31
            n_p = (rrp3 * rr41) / (rrp4 * rr31)
            n_m = (rrm3 * rr41) / (rrm4 * rr31)
            Pi_1 = (2.0 / root) * (legendre_pi(n1, am, k2) - legendre_pi(n1, aux, k2))
            Pi_p = (2.0 / root) * (rr43 / (rrp3 * rrp4)) * (legendre_pi(n_p, am, k2) - legendre_pi(n_p, aux, k2))
            Pi_m = (2.0 / root) * (rr43 / (rrm3 * rrm4)) * (legendre_pi(n_m, am, k2) - legendre_pi(n_m, aux, k2))
            I_1[:, m2] = q3 * (-t) + rr43 * Pi_1
            I_2[:, m2] = H - 0.5 * (q1 * q4 + q2 * q3) * (-t) - E
            I_p[:, m2] = t / rrp3 - Pi_p
            I_m[:, m2] = t / rrm3 - Pi_m

    I_0 = -tau                                                       # :627-631
    I_phi = (2 * a / (rplus - rminus)) * ((rplus - 0.5 * a * lam) * I_p - (rminus - 0.5 * a * lam) * I_m)
    I_tA = (4 / (rplus - rminus)) * ((rplus ** 2 - 0.5 * a * lam * rplus) * I_p
                                       - (rminus ** 2 - 0.5 * a * lam * rminus) * I_m)
    I_t = I_tA + 4 * I_0 + 2 * I_1 + I_2
    return r_s, I_phi, I_t, I_2


def coords_at_tau(a, observer_coords, image_coords, tau, full=True):
    """(sigma, [t, r, theta, phi]) at Mino times tau (kerr_raytracing_ana.py:126-193)."""
    _, r_o, th_o, _ = observer_coords
    alpha, beta = _arr(image_coords[0]), _arr(image_coords[1])
    tau = np.asarray(tau, dtype=float)
    if tau.ndim == 1:
        tau = tau.reshape(1, -1)
    lam, eta = conserved(a, th_o, alpha, beta)
    if abs(a) < MINSPIN and np.any(eta < 0):
        eta = np.where(eta < 0, EP, eta)
    s_o = sign_pos0(beta)
    up, um, _, _ = angular_consts(a, th_o, lam, eta)
    r1, r2, r3, r4, _ = quartic_roots(a, lam, eta)
    th_s, G_ph, G_t = theta_motion(a, th_o, s_o, lam, eta, up, um, tau, full)
    r_s, I_ph, I_t, I_sig = radial_motion(a, r_o, lam, eta, r1, r2, r3, r4, tau, full)
    sig = I_sig + a ** 2 * G_t
    t_s = I_t + a ** 2 * G_t
    ph_s = I_ph + lam * G_ph
    return sig, [t_s, r_s, th_s, ph_s]


def raytrace(a, observer_coords, image_coords, ngeo, maxtaufrac=1.0 - 1.0e-10):
    """tausteps, sigma, [t, r, theta, phi] of raytrace_ana (kerr_raytracing_ana.py:46-122)."""
    _, r_o, th_o, _ = observer_coords
    alpha, beta = _arr(image_coords[0]), _arr(image_coords[1])
    lam, eta = conserved(a, th_o, alpha, beta)
    if abs(a) < MINSPIN and np.any(eta < 0):
        eta = np.where(eta < 0, EP, eta)
    r1, r2, r3, r4, _ = quartic_roots(a, lam, eta)
    tau_tot = mino_total(a, r_o, eta, r1, r2, r3, r4)
    tausteps = np.linspace(0, maxtaufrac * tau_tot, ngeo)
    sig, coords = coords_at_tau(a, observer_coords, [alpha, beta], tausteps, True)
    return tausteps, sig, coords


# --------------------------------------------------------------------------------------
# crossing counts and equatorial lensing
# --------------------------------------------------------------------------------------
def outside_crit(a, th_o, alpha, beta):
    """1 where (alpha,beta) is outside the critical curve (kerr_raytracing_utils.py:794-826)."""
    rh = 1 + np.sqrt(1 - a ** 2)
    lam, eta = conserved(a, th_o, alpha, beta)
    _, _, _, r4, _ = quartic_roots(a, lam, eta)
    return ((r4.imag == 0) & (r4.real > rh) & ~(eta < 0)).astype(float)


def n_poloidal(a, th_o, alpha, beta, tau):
    """Fractional number of poloidal orbits (kerr_raytracing_utils.py:551-603)."""
    lam, eta = conserved(a, th_o, alpha, beta)
    out = np.full(np.shape(tau), -2.0)
    reg = ~(eta <= 0)
    if reg.any():
        _, _, ratio, a2um = angular_consts(a, th_o, lam[reg], eta[reg])
        K = sp.ellipk(ratio)
        out[..., reg] = (np.sqrt(-a2um) * np.asarray(tau)[..., reg]) / (4 * K)
    return out


def _crossing_terms(a, th_o, lam, eta):
    up, um, ratio, a2um = angular_consts(a, th_o, lam, eta)
    F0 = sp.ellipkinc(np.arcsin(np.cos(th_o) / np.sqrt(up)), ratio)
    K = sp.ellipkinc(0.5 * np.pi, ratio)
    return F0, K, np.sqrt(-a2um)


def n_equatorial(a, th_o, alpha, beta, tau):
    """Number of equatorial crossings up to tau (kerr_raytracing_utils.py:605-684)."""
    lam, eta = conserved(a, th_o, alpha, beta)
    tau = np.asarray(tau, dtype=float)
    out = np.full(tau.shape, -2.0)
    reg = ~(eta <= 0)
    if reg.any():
        F0, K, w = _crossing_terms(a, th_o, lam[reg], eta[reg])
        b = beta[reg]
        t = tau[..., reg]
        if th_o < np.pi / 2:
            first = b <= 0
            val = np.where(first, np.floor((t * w - F0) / (2 * K)), np.floor((t * w + F0) / (2 * K)) - 1)
        else:
            first = b >= 0
            val = np.where(first, np.floor((t * w + F0) / (2 * K)), np.floor((t * w - F0) / (2 * K)) - 1)
        out[..., reg] = val
    return out


def n_angular_turnings(a, th_o, alpha, beta, tau):
    """Fractional number of angular turnings at Mino time tau (kerr_raytracing_utils.py:686-792): ordinary rays
    (tau w -+ F0)/(2K) + 1/2 with the sign by beta <= 0 / > 0; the vortical branch with its own amplitude, parameter 1 - u+/u-
    and the reference's extra half turn once the first one is complete (beta <= 0)."""
    lam, eta = conserved(a, th_o, alpha, beta)
    tau = np.asarray(tau, dtype=float)
    out = np.empty(tau.shape)
    vort = eta <= 0
    reg = ~vort
    if reg.any():
        F0, K, w = _crossing_terms(a, th_o, lam[reg], eta[reg])
        sgn = np.where(beta[reg] <= 0, -1.0, 1.0)
        out[..., reg] = (tau[..., reg] * w + sgn * F0) / (2 * K) + 0.5
    if vort.any():
        up, um, ratio, a2um = angular_consts(a, th_o, lam[vort], eta[vort])
        nu = np.arcsin(np.sqrt((np.cos(th_o) ** 2 - um) / (up - um)))
        F0 = sp.ellipkinc(nu, 1 - ratio)
        K = sp.ellipkinc(0.5 * np.pi, 1 - ratio)
        w = np.sqrt(a2um)
        neg = beta[vort] <= 0
        first = (tau[..., vort] * w - F0) / (2 * K) + 0.5
        first = np.where(first > 0.5, first + 0.5, first)
        out[..., vort] = np.where(neg, first, (tau[..., vort] * w + F0) / (2 * K) + 0.5)
    return out


def r_equatorial(a, r_o, th_o, mbar, alpha, beta):
    """(r_s, Ir, Imax, Nmax) of the mbar-th equatorial crossing (equatorial_lensing.py:85-196)."""
    alpha, beta = _arr(alpha), _arr(beta)
    lam, eta = conserved(a, th_o, alpha, beta)
    r_s = np.full(alpha.shape, -1.0)
    Ir = np.full(alpha.shape, -1.0)
    Imax = np.full(alpha.shape, -1.0)
    Nmax = np.full(alpha.shape, -2.0)
    reg = ~(eta <= 0)
    if reg.any():
        lr, er, b = lam[reg], eta[reg], beta[reg]
        F0, K, w = _crossing_terms(a, th_o, lr, er)
        r1, r2, r3, r4, _ = quartic_roots(a, lr, er)
        tmax = mino_total(a, r_o, er, r1, r2, r3, r4)
        if th_o < np.pi / 2:                                           # :158-169
            m = mbar + np.heaviside(b, 0)
            first = b <= 0
            nmax = np.where(first, np.floor((tmax * w - F0) / (2 * K)), np.floor((tmax * w + F0) / (2 * K)) - 1)
            ir = np.where(first, (2 * m * K + F0) / w, (2 * m * K - F0) / w)
        else:                                                          # :170-180
            m = mbar + np.heaviside(-b, 0)
            first = b >= 0
            nmax = np.where(first, np.floor((tmax * w + F0) / (2 * K)), np.floor((tmax * w - F0) / (2 * K)) - 1)
            ir = np.where(first, (2 * m * K - F0) / w, (2 * m * K + F0) / w)
        rs, _, _, _ = radial_motion(a, r_o, lr, er, r1, r2, r3, r4, ir.reshape(1, -1), full=False)
        rs = rs[0]
        rs[nmax < mbar] = 0
        r_s[reg], Ir[reg], Imax[reg], Nmax[reg] = rs, ir, tmax, nmax
    return r_s, Ir, Imax, Nmax


INF_LENS = 1.0e50     # kgeo/equatorial_lensing.py:18
R0_LENS = np.inf      # kgeo/equatorial_lensing.py:19


def r_equatorial2(rho, varphi, a, th0, mbar=0):
    """Crossing radius along a polar ray of the image plane, made sign-definite for the root find
    (equatorial_lensing.py:197-217)."""
    alpha = np.cos(varphi) * rho
    beta = np.sin(varphi) * rho
    r = r_equatorial(a, R0_LENS, th0, mbar, np.array([alpha]), np.array([beta]))[0][0]
    oc = outside_crit(a, th0, np.array([alpha]), np.array([beta]))[0]
    if r <= 0 and oc == 1:
        r = INF_LENS
    elif r <= 0:
        r = 1
    return r


def _rho_bracket(req, mbar):
    """equatorial_lensing.py:232-241 (bounding ranges eyeballed from Gralla+Lupsasca Fig 6)"""
    if mbar == 0:
        return 0, req + 5
    if mbar == 1:
        return 0, 100
    return 1, 10


def rho_of_req_single(a, th0, req, varphi, mbar=0):
    """equatorial_lensing.py:249-266: scipy.optimize.brentq (default xtol=2e-12, rtol=4 eps, maxiter=100) on
    r_equatorial2(rho) - req."""
    from scipy.optimize import brentq
    rhomin, rhomax = _rho_bracket(req, mbar)
    return brentq(lambda rho: r_equatorial2(rho, varphi, a, th0, mbar) - req, rhomin, rhomax)


def rho_of_req(a, th0, req, mbar=0, varphis=None):
    """equatorial_lensing.py:226-247"""
    if varphis is None:
        varphis = np.linspace(-180, 179, 360) * np.pi / 180.
    rhos = np.array([rho_of_req_single(a, th0, req, v, mbar) for v in varphis])
    return varphis, rhos, np.cos(varphis) * rhos, np.sin(varphis) * rhos


def critical_curve(a, th0, n=100000):
    """equatorial_lensing.py:268-298 (GL19 eqs 38-40)"""
    rminus = 2 * (1 + np.cos(2. * np.arccos(-a) / 3.))
    rplus = 2 * (1 + np.cos(2. * np.arccos(a) / 3.))
    r = np.linspace(rminus, rplus, n)
    lam = -(r ** 3 + a ** 2 + r * (a ** 2) - 3 * r ** 2) / (a * r - a)
    eta = (r ** 2) * (3 * r ** 2 + a ** 2 - lam ** 2) / (r ** 2 - a ** 2)
    alpha = -lam / np.sin(th0)
    beta = np.sqrt(eta + (a * np.cos(th0)) ** 2 - (lam / np.tan(th0)) ** 2)
    mask = np.isnan(alpha) + np.isnan(beta)
    alpha = alpha[~mask]
    beta = beta[~mask]
    return np.hstack((alpha, np.flip(alpha))), np.hstack((beta, -np.flip(beta)))


def nmax_equatorial(a, r_o, th_o, alpha, beta):
    """equatorial_lensing.py:22-51"""
    alpha, beta = _arr(alpha), _arr(beta)
    lam, eta = conserved(a, th_o, alpha, beta)
    r1, r2, r3, r4, _ = quartic_roots(a, lam, eta)
    return n_equatorial(a, th_o, alpha, beta, mino_total(a, r_o, eta, r1, r2, r3, r4))


def nmax_poloidal(a, r_o, th_o, alpha, beta):
    """equatorial_lensing.py:53-82"""
    alpha, beta = _arr(alpha), _arr(beta)
    lam, eta = conserved(a, th_o, alpha, beta)
    r1, r2, r3, r4, _ = quartic_roots(a, lam, eta)
    return n_poloidal(a, th_o, alpha, beta, mino_total(a, r_o, eta, r1, r2, r3, r4))


# --------------------------------------------------------------------------------------
# physics closed forms in the equatorial plane
# models are plain dicts: {'type': str, **params}
# --------------------------------------------------------------------------------------
def _isco(a, s):
    z1 = 1 + np.cbrt(1 - a ** 2) * (np.cbrt(1 + a) + np.cbrt(1 - a))
    z2 = np.sqrt(3 * a ** 2 + z1 ** 2)
    return 3 + z2 - s * np.sqrt((3 - z1) * (3 + z1 + 2 * z2))


def u_zamo(a, r):
    """kgeo/velocities.py:179-208 (theta = pi/2)."""
    th = np.pi / 2.
    Delta = r ** 2 - 2 * r + a ** 2
    Sigma = r ** 2 + a ** 2 * np.cos(th) ** 2
    g00 = -(1 - 2 * r / Sigma)
    g33 = (r ** 2 + a ** 2 + 2 * r * (a * np.sin(th)) ** 2 / Sigma) * np.sin(th) ** 2
    g03 = -2 * r * a * np.sin(th) ** 2 / Sigma
    v3 = -g03 / g33
    u0 = np.sqrt(-1. / (g00 + 2 * g03 * v3 + g33 * v3 * v3))
    return u0, 0 * r, 0 * r, u0 * v3


def u_infall(a, r):
    """kgeo/velocities.py:211-226."""
    Delta = r ** 2 + a ** 2 - 2 * r
    Xi = (r ** 2 + a ** 2) ** 2 - Delta * a ** 2
    return Xi / (r ** 2 * Delta), -np.sqrt(2 * r * (r ** 2 + a ** 2)) / (r ** 2), 0 * r, 2 * a / (r * Delta)


def u_kep(a, r, retrograde=False):
    """kgeo/velocities.py:229-281."""
    s = -1 if retrograde else 1
    ri = _isco(a, s)
    spin, asign = np.abs(a), np.sign(a)
    u0, u1, u3 = np.zeros(r.shape), np.zeros(r.shape), np.zeros(r.shape)
    out = r >= ri
    if out.any():
        rr = r[out]
        u0[out] = (rr ** 1.5 + s * spin) / np.sqrt(rr ** 3 - 3 * rr ** 2 + 2 * s * spin * rr ** 1.5)
        u3[out] = (asign * s / (rr ** 1.5 + s * spin)) * u0[out]
    if (~out).any():
        rr = r[~out]
        Delta = rr ** 2 - 2 * rr + a ** 2
        ell_i = s * asign * (ri ** 2 + a ** 2 - s * 2 * spin * np.sqrt(ri)) / (ri ** 1.5 - 2 * np.sqrt(ri) + s * spin)
        gam_i = np.sqrt(1 - 2. / (3. * ri))
        H = (2 * rr - a * ell_i) / Delta
        u0[~out] = gam_i * (1 + (2 / rr) * (1 + H))
        u1[~out] = -np.sqrt(2. / (3 * ri)) * (ri / rr - 1) ** 1.5
        u3[~out] = gam_i * (ell_i + a * H) / (rr ** 2)
    return u0, u1, 0 * r, u3


def u_subkep(a, r, fac_subkep=1, retrograde=False):
    """kgeo/velocities.py:284-359."""
    s = -1 if retrograde else 1
    ri = _isco(a, s)
    spin, asign = np.abs(a), np.sign(a)
    u0, u1, u3 = np.zeros(r.shape), np.zeros(r.shape), np.zeros(r.shape)
    out = r >= ri
    if out.any():
        rr = r[out]
        Delta = rr ** 2 - 2 * rr + a ** 2
        Xi = (rr ** 2 + a ** 2) ** 2 - Delta * a ** 2
        ell = asign * s * (rr ** 2 + spin ** 2 - s * 2 * spin * np.sqrt(rr)) / (rr ** 1.5 - 2 * np.sqrt(rr) + s * spin)
        ell = ell * fac_subkep
        gam = np.sqrt(Delta / (Xi / rr ** 2 - 4 * a * ell / rr - (1 - 2 / rr) * ell ** 2))
        H = (2 * rr - a * ell) / Delta
        chi = 1 / (1 + (2 / rr) * (1 + H))
        Omega = (chi / rr ** 2) * (ell + a * H)
        u0[out] = gam / chi
        u3[out] = (gam / chi) * Omega
    if (~out).any():
        rr = r[~out]
        Delta = rr ** 2 - 2 * rr + a ** 2
        Xi = (rr ** 2 + a ** 2) ** 2 - Delta * a ** 2
        Delta_i = ri ** 2 - 2 * ri + a ** 2
        Xi_i = (ri ** 2 + a ** 2) ** 2 - Delta_i * a ** 2
        ell_i = asign * s * (ri ** 2 + spin ** 2 - s * 2 * spin * np.sqrt(ri)) / (ri ** 1.5 - 2 * np.sqrt(ri) + s * spin)
        ell_i = ell_i * fac_subkep
        gam_i = np.sqrt(Delta_i / (Xi_i / ri ** 2 - 4 * a * ell_i / ri - (1 - 2 / ri) * ell_i ** 2))
        H = (2 * rr - a * ell_i) / Delta
        chi = 1 / (1 + (2 / rr) * (1 + H))
        Omega = (chi / rr ** 2) * (ell_i + a * H)
        nu = (rr / Delta) * np.sqrt(Xi / rr ** 2 - 4 * a * ell_i / rr - (1 - 2 / rr) * ell_i ** 2 - Delta / gam_i ** 2)
        u0[~out] = gam_i / chi
        u1[~out] = -gam_i * (Delta / rr ** 2) * nu
        u3[~out] = (gam_i / chi) * Omega
    return u0, u1, 0 * r, u3


def u_gelles(a, r, beta=0.3, chi=-150 * (np.pi / 180.)):
    """kgeo/velocities.py:362-389."""
    a2, r2 = a ** 2, r ** 2
    Delta = r2 - 2 * r + a2
    Xi = (r2 + a2) ** 2 - Delta * a2
    omegaz = 2 * a * r / Xi
    gamma = 1 / np.sqrt(1 - beta ** 2)
    u0 = (gamma / r) * np.sqrt(Xi / Delta)
    u1 = (beta * gamma * np.cos(chi) / r) * np.sqrt(Delta)
    u3 = (gamma * omegaz / r) * np.sqrt(Xi / Delta) + (r * beta * gamma * np.sin(chi)) / np.sqrt(Xi)
    return u0, u1, 0 * r, u3


def u_simfit(a, r, ell_isco=1., vr_isco=2., p1=6., p2=2., dd=0.2):
    """kgeo/velocities.py:392-439."""
    r_isco = _isco(a, 1)
    a2, r2 = a ** 2, r ** 2
    Delta = r2 - 2 * r + a2
    Sigma = r2
    g00_up = -(r2 + a2 + 2 * r * a2 / Sigma) / Delta
    g11_up = Delta / Sigma
    g33_up = (Delta - a2) / (Sigma * Delta)
    g03_up = -(2 * r * a) / (Sigma * Delta)
    ell = ell_isco * (r / r_isco) ** .5
    vr = -vr_isco * ((r / r_isco) ** (-p1)) * (0.5 * (1 + (r / r_isco) ** (1 / dd))) ** ((p1 - p2) * dd)
    gam = np.sqrt(-1. / (g00_up + g11_up * vr * vr + g33_up * ell * ell - 2 * g03_up * ell))
    u_0, u_1, u_3 = -gam, vr * gam, ell * gam
    return g00_up * u_0 + g03_up * u_3, g11_up * u_1, 0 * r, g33_up * u_3 + g03_up * u_0


def u_general(a, r, fac_subkep=1, beta_phi=1, beta_r=1, retrograde=False):
    """kgeo/velocities.py:442-470."""
    Delta = r ** 2 + a ** 2 - 2 * r
    u0i, u1i, _, u3i = u_infall(a, r)
    u0s, u1s, _, u3s = u_subkep(a, r, retrograde=retrograde, fac_subkep=fac_subkep)
    Oi, Os = u3i / u0i, u3s / u0s
    u1 = u1s + (1 - beta_r) * (u1i - u1s)
    Omega = Os + (1 - beta_phi) * (Oi - Os)
    u0 = np.sqrt(1 + (r ** 2) * (u1 ** 2) / Delta)
    u0 = u0 / np.sqrt(1 - (r ** 2 + a ** 2) * Omega ** 2 - (2 / r) * (1 - a * Omega) ** 2)
    return u0, u1, 0 * r, u0 * Omega


def velocity_u_lab(vel, a, r, th=np.pi / 2):
    """Dispatch of Velocity.u_lab (kgeo/velocities.py:70-96); th is only honoured by the drift frame (jet models)."""
    r = _arr(r)
    t = vel['type']
    if t == 'zamo':
        return u_zamo(a, r)
    if t == 'infall':
        return u_infall(a, r)
    if t in ('kep', 'cunningham'):
        return u_kep(a, r, retrograde=vel.get('retrograde', False))
    if t in ('subkep', 'cunningham_subkep'):
        return u_subkep(a, r, retrograde=vel.get('retrograde', False), fac_subkep=vel.get('fac_subkep', 1))
    if t == 'general':
        return u_general(a, r, retrograde=vel.get('retrograde', False), fac_subkep=vel.get('fac_subkep', 1),
                         beta_phi=vel.get('beta_phi', 1), beta_r=vel.get('beta_r', 1))
    if t == 'gelles':
        return u_gelles(a, r, beta=vel.get('gelles_beta', 0.3), chi=vel.get('gelles_chi', -150 * np.pi / 180.))
    if t == 'simfit':
        return u_simfit(a, r, vel.get('ell_isco', 1.), vel.get('vr_isco', 2.), vel.get('p1', 6.), vel.get('p2', 2.),
                        vel.get('dd', 0.2))
    if t == 'fromfile':
        return u_from_table(a, r, load_table(vel['file'], ('U1', 'U2', 'U3')))
    if t == 'driftframe':
        return u_driftframe(a, r, vel['bfield'], vel.get('nu_parallel', 0), th, vel.get('gammamax'))
    raise Exception("veltype %s not in oracle" % t)


def load_table(filename, columns):
    """CSV reader of the 'fromfile' models (kgeo/velocities.py:626-657, kgeo/bfields.py:749-787)."""
    with open(filename, 'r') as f:
        header = [x.strip() for x in f.readline().strip()[1:].split(',')]
    data = np.loadtxt(filename, delimiter=',', skiprows=1)
    radii = data[:, header.index('r')]
    return [radii] + [data[:, header.index(c)] if c in header else np.zeros_like(radii) for c in columns]


def u_from_table(a, r, tab):
    """u_from_u123 (kgeo/velocities.py:659-698): np.interp of the primitives, then normalisation."""
    radii, U1, U2, U3 = tab
    u1 = np.interp(r, radii, U1)
    u2 = np.interp(r, radii, U2)
    u3 = np.interp(r, radii, U3)
    r2 = r * r
    a2 = a * a
    th = np.pi / 2.
    Delta = r2 - 2 * r + a2
    Sigma = r2 + a2 * np.cos(th) ** 2
    g11 = Sigma / Delta
    g22 = Sigma
    g33 = (r2 + a2 + 2 * r * (a * np.sin(th)) ** 2 / Sigma) * np.sin(th) ** 2
    gcon00 = - (r2 + a2 + 2 * r * a2 / Sigma * np.sin(th) ** 2) / Delta
    gcon03 = -2 * r * a / Sigma / Delta
    alpha = 1. / np.sqrt(-gcon00)
    gamma = np.sqrt(1. + g11 * u1 * u1 + g22 * u2 * u2 + g33 * u3 * u3)
    return gamma / alpha, u1, u2, u3 - gamma * alpha * gcon03


def _metric_eq(a, r, th=np.pi / 2):
    """Covariant BL metric, by default at theta=pi/2 with the reference's cos(pi/2)=6e-17 kept
    (kgeo/velocities.py:102-114)."""
    a2, r2 = a ** 2, r ** 2
    cth2, sth2 = np.cos(th) ** 2, np.sin(th) ** 2
    Delta = r2 - 2 * r + a2
    Sigma = r2 + a2 * cth2
    g00 = -(1 - 2 * r / Sigma)
    g11 = Sigma / Delta
    g22 = Sigma
    g33 = (r2 + a2 + 2 * r * a2 * sth2 / Sigma) * sth2
    g03 = -2 * r * a * sth2 / Sigma
    return Delta, Sigma, g00, g11, g22, g33, g03


def velocity_u_cov(vel, a, r, th=np.pi / 2):
    """kgeo/velocities.py:98-122."""
    u0, u1, u2, u3 = velocity_u_lab(vel, a, r, th)
    _, _, g00, g11, g22, g33, g03 = _metric_eq(a, r, th)
    return g00 * u0 + g03 * u3, g11 * u1, g22 * u2, g33 * u3 + g03 * u0


def velocity_tetrads(vel, a, r, th=np.pi / 2):
    """kgeo/velocities.py:124-176."""
    sth2 = np.sin(th) ** 2
    Delta, Sigma, g00, g11, g22, g33, g03 = _metric_eq(a, r, th)
    u0, u1, u2, u3 = velocity_u_lab(vel, a, r, th)
    l0, l1, l2, l3 = velocity_u_cov(vel, a, r, th)
    Nr = np.sqrt(-g11 * (l0 * u0 + l3 * u3) * (1 + l2 * u2))
    Nth = np.sqrt(g22 * (1 + l2 * u2))
    Nph = np.sqrt(-Delta * sth2 * (l0 * u0 + l3 * u3))
    et = (-u0, -u1, -u2, -u3)
    ex = (l1 * u0 / Nr, -(l0 * u0 + l3 * u3) / Nr, 0, l1 * u3 / Nr)
    ey = (l2 * u0 / Nth, l2 * u1 / Nth, (1 + l2 * u2) / Nth, l2 * u3 / Nth)
    ez = (l3 / Nph, 0, 0, -l0 / Nph)
    return et, ex, ey, ez


def bfield_lab(bf, a, r):
    """Lab-frame B^i at theta=pi/2 (kgeo/bfields.py:93-131 + model functions)."""
    r = _arr(r)
    t = bf['type']
    th = np.pi / 2
    if t == 'fromfile':                                            # bfields.py:790-806
        radii, Br, Bth, Bph = load_table(bf['file'], ('Br', 'Btheta', 'Bphi'))
        return np.interp(r, radii, Br), np.interp(r, radii, Bth), np.interp(r, radii, Bph)
    if t in ('simple', 'rad', 'vert', 'tor', 'simple_rm1'):
        if t == 'rad':
            cr, cv, cp = 1, 0, 0
        elif t == 'vert':
            cr, cv, cp = 0, 1, 0
        elif t == 'tor':
            cr, cv, cp = 0, 0, 1
        else:
            cr, cv, cp = bf.get('Cr', 0), bf.get('Cvert', 0), bf.get('Cph', 0)
        Sigma = r ** 2 + a ** 2 * np.cos(th) ** 2
        gdet = np.abs(np.sin(th)) * Sigma
        rad = (r * np.sin(th) / gdet) if t == 'simple_rm1' else (np.sin(th) / gdet)     # bfields.py:447 / :473
        return cr * rad + cv * (r * np.cos(th) / gdet), cv * (-np.sin(th) / gdet), cp * (1. / gdet)
    if t == 'bz_guess':                                                                    # bfields.py:504-542
        C = bf.get('C', 1)
        a2, r2 = a ** 2, r ** 2
        sth2 = np.sin(th) ** 2
        Delta = r2 - 2 * r + a2
        Pi = (r2 + a2) ** 2 - a2 * Delta * sth2
        rH = 1 + np.sqrt(1 - a2)
        OmegaH = a / (2 * rH)
        spin = np.abs(a)
        f2 = (6 * np.pi * np.pi - 49.) / 72.
        omega2 = 1. / 32. - (4 * f2 - 1) * sth2 / 64.
        OmegaBZ = spin / 8. + (spin ** 3) * omega2
        if a < 0:
            OmegaBZ *= -1
        ratio = -1 * (OmegaH - OmegaBZ) * np.sqrt(Pi) / Delta
        Br = C / r2
        return Br, 0 * Br, ratio * Br
    if t == 'gen_power':                                                                   # bfields.py:712-746
        n_I, p_val, above = bf.get('n_I', 0), bf.get('p_val', 1), bf.get('isAbove', True)
        a2, r2 = a ** 2, r ** 2
        rp = 1 + np.sqrt(1 - a2)
        sth, cth = np.sin(th), np.cos(th)
        Delta = r2 - 2 * r + a2
        Sigma = r2 + a2 * cth ** 2
        gdet = sth * Sigma
        sg = 1.0 if above else -1.0
        I_0 = sg * (a * np.pi * sth ** 2) / (rp * rp + a2 * cth ** 2)
        dpsidtheta = sg * sth * ((r / rp) ** p_val)
        I = I_0 * (r / rp) ** (-1 * n_I)
        dpsidr = (1 - np.abs(cth)) * ((p_val * (r ** (p_val - 1))) / (rp ** p_val))
        return dpsidtheta / gdet, -1 * (dpsidr / gdet), I / (2 * np.pi * Delta * sth ** 2)
    if t == 'monopoleA':                                                                   # bfields.py:480-501
        C, omegafac = bf.get('C', 1), bf.get('omegafac', 0.5)
        rh = 1 + np.sqrt(1 - a ** 2)
        omegah = a / (2 * rh)
        phih = 4 * np.pi * C
        cth = np.cos(th)
        Delta = r ** 2 - 2 * r + a ** 2
        Sigma = r ** 2 + a ** 2 * cth ** 2
        omegaf = omegafac * omegah
        I = phih * (np.sin(th) ** 2) * rh * (omegaf - omegah) / (rh ** 2 + a ** 2 * cth ** 2)
        return (phih / (4 * np.pi)) / Sigma, 0 * r, I / (2 * np.pi * Delta * (np.sin(th) ** 2))
    if t in ('bz_monopole', 'bz_para', 'power'):
        return bfield_lab_omega(bf, a, r)[:3]
    raise Exception("fieldtype %s not in oracle" % t)


_BZ_FR = None


def bz_fr_table():
    """The reference's bz_fr_data.dat rebuilt from its recipe (kgeo/bfields.py:7-20): f(r) on linspace(1, 100, 100000); real
    part of the complex formula, Li2(x) = spence(1 - x) and Re Li2(x) = pi^2/3 - ln^2(x)/2 - Li2(1/x) for x > 1."""
    global _BZ_FR
    if _BZ_FR is None:
        r = np.linspace(1.0, 100.0, 100000)
        x = 2.0 / r
        with np.errstate(all="ignore"):
            li = np.where(x <= 1, sp.spence(1 - np.minimum(x, 1.0)),
                          np.pi ** 2 / 3.0 - 0.5 * np.log(x) ** 2 - sp.spence(1 - np.minimum(1.0 / x, 1.0)))
            f = (li - np.log(np.abs(1 - x)) * np.log(r / 2.0)) * r * r * (2 * r - 3) / 8.0
            f += (1 + 3 * r - 6 * r * r) * np.log(r / 2.0) / 12.0
            f += 11.0 / 72.0 + 1.0 / (3.0 * r) + r / 2.0 - r * r / 2.0
        _BZ_FR = (r, f)
    return _BZ_FR


def _linear_spline(xp, fp, x):
    """UnivariateSpline(k=1, s=0, ext=0) and its derivative (kgeo/bfields.py:27-30): linear interpolation, linear
    extrapolation with the end segments, slope of the segment holding x."""
    i = np.clip(np.searchsorted(xp, x, side='right') - 1, 0, len(xp) - 2)
    h = xp[i + 1] - xp[i]
    return (fp[i] * (xp[i + 1] - x) + fp[i + 1] * (x - xp[i])) / h, (fp[i + 1] - fp[i]) / h


def bfield_lab_omega(bf, a, r, th=np.pi / 2):
    """(B1, B2, B3, omega) of the models with a field-line angular speed (kgeo/bfields.py:175-197 and the model functions);
    by default theta = pi/2 with the reference's cos(pi/2) = 6e-17; the jet models pass the polar angle of the crossing."""
    r = _arr(r)
    t = bf['type']
    a2, r2 = a ** 2, r ** 2
    sth, cth = np.sin(th), np.cos(th)
    abscth, cth2, sth2 = np.abs(cth), cth ** 2, sth ** 2
    Delta = r2 - 2 * r + a2
    Sigma = r2 + a2 * cth2
    gdet = sth * Sigma
    rp = 1 + np.sqrt(1 - a2)
    if t == 'bz_monopole':                                                                # bfields.py:545-592
        C, second = bf.get('C', 1), bf.get('secondorder_only', False)
        xs, fs = bf['fr_table'] if bf.get('fr_table') is not None else bz_fr_table()
        fr, frp = _linear_spline(xs, fs, r)
        far = r > np.max(xs)
        fr = np.where(far, 1. / (4. * r), fr)
        frp = np.where(far, -1. / (4. * r2), frp)
        dphidtheta = C * sth * (1 + 0.5 * a2 * fr * (1 + 3 * np.cos(2 * th)))
        dphidr = C * a2 * sth2 * cth * frp
        f2 = (6 * np.pi * np.pi - 49.) / 72.
        omega2 = 1. / 32. - sth2 * (4 * f2 - 1) / 64.
        I = -1 * C * 2 * np.pi * sth2 * (a / 8. + (a ** 3) * (omega2 + 0.25 * fr * cth2))
        Br, Bth, Bph = dphidtheta / gdet, -dphidr / gdet, I / (2 * np.pi * Delta * sth2)
        if second:
            Br = C / r2 + C * a2 * (-cth2 / (r2 * r2) + fr * (0.5 + 1.5 * np.cos(2 * th)) / r2)
            Bth = -C * a2 * frp * cth * sth / r2
            Bph = -C * 0.125 * a / (r2 - 2 * r) + C * (a ** 3) * (0.125 / ((r2 - 2 * r) ** 2) - (omega2 + 0.25 * cth2 * fr) / ((r2 - 2 * r)))
        spin = np.abs(a)
        Om = spin / 8. + spin ** 3 * omega2
        if a < 0:
            Om *= -1
        return Br, Bth, Bph, Om + 0 * r
    if t == 'bz_para':                                                                    # bfields.py:609-673
        C = bf.get('C', 1)
        psi = r * (1 - abscth) + rp * (1 + abscth) * (1 - np.log(1 + abscth)) - 2 * rp * (1 - np.log(2))
        dpsidtheta = np.sign(cth) * sth * (r + rp * np.log(1 + abscth))
        dpsidr = 1 - abscth
        argvals = -psi / rp + np.log(4)
        ok = argvals >= -1 / np.e
        w = np.real(sp.lambertw(np.where(ok, argvals, 0.0), k=0))
        x = np.exp(w)
        y = 1 + psi / (1 + w) / (x * (2 - x))
        Om = np.where(ok, C * (a / 4 / y), 0.0)        # the reference's complex value times the mask, real part
        I = -4 * np.pi * psi * Om * np.sign(cth)
        return C * dpsidtheta / gdet, -C * dpsidr / gdet, I / (2 * np.pi * Delta * sth2), Om
    if t == 'power':                                                                      # bfields.py:618-625,676-709
        p, C, usemono = bf.get('p', 1), bf.get('C', 1), bf.get('usemono', False)
        psi = r ** p * (1 - abscth)
        dpsidtheta = np.sign(cth) * sth * (r ** p)
        dpsidr = p * psi / r
        if usemono or p == 0:
            Om = a / 8 + 0 * r
        else:
            cthh = 1 - psi * rp ** (-p)
            Om = a / (4 + 8 / (1 + cthh))
        I = -4 * np.pi * psi * Om * np.sign(cth) if p > 0 else -2 * np.pi * psi * (2 - psi) * Om * np.sign(cth)
        return C * dpsidtheta / gdet, -C * dpsidr / gdet, I / (2 * np.pi * Delta * sth2), Om
    if np.any(np.asarray(th) != np.pi / 2):
        raise Exception("fieldtype %s: only the equatorial plane is restated" % t)
    B1, B2, B3 = bfield_lab(bf, a, r)
    if t == 'bz_guess':
        spin = np.abs(a)
        f2 = (6 * np.pi * np.pi - 49.) / 72.
        Om = spin / 8. + (spin ** 3) * (1. / 32. - (4 * f2 - 1) * sth2 / 64.)
        return B1, B2, B3, (-Om if a < 0 else Om) + 0 * r
    if t == 'monopoleA':
        return B1, B2, B3, bf.get('omegafac', 0.5) * a / (2 * rp) + 0 * r
    if t == 'gen_power':
        return B1, B2, B3, 0.5 * a / (2 * rp) + 0 * r
    raise Exception("self.omega_field not implemented for fieldtype %s'!" % t)


def get_eco(r0, theta, Omegaf, spin, M=1.0):
    """E - Omega_F L as a function of the launch point (kgeo/ff_boost.py:102-119 with metric / invmetric of :122-146)."""
    Sig = r0 ** 2 + spin ** 2 * np.cos(theta) ** 2
    Del = r0 ** 2 - 2 * M * r0 + spin ** 2
    s2 = np.sin(theta) ** 2
    gtt = -(1 - 2 * M * r0 / Sig)
    gtphi = -2 * spin * M * r0 / Sig * s2
    gphiphi = (r0 ** 2 + spin ** 2 + 2 * M * r0 * spin ** 2 / Sig * s2) * s2
    ginvtt = -1 / Del * (r0 ** 2 + spin ** 2 + 2 * M * r0 * spin ** 2 / Sig * s2)
    ginvtphi = -2 * M * r0 * spin / (Sig * Del)
    ginvphiphi = (Del - spin ** 2 * s2) / (Sig * Del * s2)
    gtphifac = gtphi + gphiphi * Omegaf
    gttfac = gtt + gtphi * Omegaf
    coef0 = (gtt + Omegaf * (2 * gtphi + gphiphi * Omegaf)) ** 2
    coef1 = ginvphiphi * gtphifac ** 2 + 2 * ginvtphi * gtphifac * gttfac + ginvtt * gttfac ** 2
    return np.sqrt(-coef0 / coef1)


def getnu_cons(bf, r, theta, r0, theta0, Omegaf, spin, M=1.0):
    """parallel boost that conserves E - Omega_F L along the field line (kgeo/velocities.py:474-486)."""
    Aconst = get_eco(r0, theta0, Omegaf, spin, M)
    r = _arr(r)
    Sig = r ** 2 + spin ** 2 * np.cos(theta) ** 2
    g03 = -2 * spin * M * r / Sig * np.sin(theta) ** 2
    g33 = (r ** 2 + spin ** 2 + 2 * M * r * spin ** 2 / Sig * np.sin(theta) ** 2) * np.sin(theta) ** 2
    alpha, vphiupper, gammap, Bhatphi = u_driftframe(spin, r, bf, 0, theta, retbunit=True)
    ffunc = gammap * (alpha - (g03 + g33 * Omegaf) * vphiupper)
    bred = Bhatphi * (g03 + g33 * Omegaf)
    nunum = ffunc * bred + np.sign(np.cos(theta)) * np.sign(r - r0) * Aconst * np.sqrt(Aconst ** 2 - ffunc ** 2 + bred ** 2)
    return np.real(nunum / (Aconst ** 2 + bred ** 2))


def u_driftframe(a, r, bf, nu_parallel=0, th=np.pi / 2, gammamax=None, retbunit=False):
    """kgeo/velocities.py:487-623 (theta = pi/2 unless given).  nu_parallel='FF' (monopole field): the energy-conserving boost
    of getnu_cons on the field line of the first crossing handed in; gammamax: the Lorentz-factor cap of :593-611 (eps < 0)."""
    r = _arr(r)
    if isinstance(nu_parallel, str):                      # :494-516
        tha = np.broadcast_to(np.asarray(th, dtype=float), r.shape)
        ind = np.where(np.nan_to_num(r) != 0)[0][0]
        omega = bfield_lab_omega(bf, a, r[ind:ind + 1], tha[ind])[3]
        omega = float(np.broadcast_to(omega, (1,))[0])
        theta0 = tha[ind]
        r0 = r0min_mono(theta0, omega, a, 1.0)
        nu_parallel = getnu_cons(bf, r, th, r0, theta0, omega, a, 1.0)
    nu_parallel = nu_parallel * np.ones_like(r)
    if np.any(np.logical_or(nu_parallel > 1, nu_parallel < -1)):
        raise Exception("nu_parallel should be in the range (-1,1)")
    a2, r2 = a ** 2, r ** 2
    cth2, sth2 = np.cos(th) ** 2, np.sin(th) ** 2
    Delta = r2 - 2 * r + a2
    Sigma = r2 + a2 * cth2
    Xi = (r2 + a2) ** 2 - Delta * a2 * sth2
    omegaz = 2 * a * r / Xi
    gdet = Sigma * np.sin(th)
    g11, g22, g33 = Sigma / Delta, Sigma, Xi * sth2 / Sigma
    alpha = np.sqrt(Delta * Sigma / Xi)
    eta3 = 2 * a * r / np.sqrt(Delta * Sigma * Xi)
    B1, B2, B3, omegaf = bfield_lab_omega(bf, a, r, th)
    E1 = (omegaf - omegaz) * Xi * np.sin(th) * B2 / Sigma
    E2 = -(omegaf - omegaz) * Xi * np.sin(th) * B1 / (Sigma * Delta)
    E3 = 0
    Bsq = g11 * B1 * B1 + g22 * B2 * B2 + g33 * B3 * B3
    Esq = g11 * E1 * E1 + g22 * E2 * E2 + g33 * E3 * E3
    B1c, B2c, B3c = g11 * B1, g22 * B2, g33 * B3
    E1c, E2c, E3c = g11 * E1, g22 * E2, g33 * E3
    vperp1 = (alpha / (Bsq * gdet)) * (E2c * B3c - B2c * E3c)
    vperp2 = (alpha / (Bsq * gdet)) * (E3c * B1c - B3c * E1c)
    vperp3 = (alpha / (Bsq * gdet)) * (E1c * B2c - B1c * E2c)
    vpar_max = np.sqrt(1 - Esq / Bsq)
    v1 = vperp1 + nu_parallel * vpar_max * B1 / np.sqrt(Bsq)
    v2 = vperp2 + nu_parallel * vpar_max * B2 / np.sqrt(Bsq)
    v3 = vperp3 + nu_parallel * vpar_max * B3 / np.sqrt(Bsq)
    if retbunit:                                          # :588-589
        return alpha, v3, 1 / vpar_max, B3 / np.sqrt(Bsq)
    vsq = g11 * v1 * v1 + g22 * v2 * v2 + g33 * v3 * v3
    gamma = 1. / np.sqrt(1 - vsq)
    if gammamax:                                          # :593-611
        pval0 = 2.0
        gmax = gammamax * np.ones_like(gamma)
        gammaeff = (1 / gmax ** pval0 + 1 / gamma ** pval0) ** (-1 / pval0)
        argdiv = np.argmin(np.abs(np.nan_to_num(gammaeff, nan=np.inf)))
        gammaeff0 = gammaeff * gamma[argdiv] / gammaeff[argdiv]
        vsqeff = 1 - 1 / gammaeff0 ** 2
        sc = np.sqrt(vsqeff / vsq)
        v1, v2, v3 = v1 * sc, v2 * sc, v3 * sc
        gamma = np.real(gammaeff0)
    return gamma / alpha, gamma * v1, gamma * v2, gamma * (v3 + eta3)


def bfield_fluid(bf, vel, a, r):
    """b^mu in the fluid frame, ideal MHD (kgeo/bfields.py:330-363)."""
    B1, B2, B3 = bfield_lab(bf, a, r)
    u0, u1, u2, u3 = velocity_u_lab(vel, a, r)
    l0, l1, l2, l3 = velocity_u_cov(vel, a, r)
    b0 = B1 * l1 + B2 * l2 + B3 * l3
    return b0, (B1 + b0 * u1) / u0, (B2 + b0 * u2) / u0, (B3 + b0 * u3) / u0


def bfield_bsq(bf, vel, a, r):
    """kgeo/bfields.py:365-396."""
    r = _arr(r)
    _, _, g00, g11, g22, g33, g03 = _metric_eq(a, r)
    b0, b1, b2, b3 = bfield_fluid(bf, vel, a, r)
    return b0 * (g00 * b0 + g03 * b3) + b1 * g11 * b1 + b2 * g22 * b2 + b3 * (g33 * b3 + g03 * b0)


_E_CGS, _ME, _C, _KB = 4.80320425e-10, 9.10938356e-28, 2.99792458e10, 1.380649e-16


def emissivity_jrest(em, a, r, g=None, sinthb=None, Bmag=None, nu_obs=230.e9, specind=0.0):
    """Emissivity.jrest (kgeo/emissivities.py:83-157)."""
    r = _arr(r)
    t = em['type']
    if t == 'thermal':
        nu_em = np.asarray(nu_obs) / np.asarray(g)
        x = r / em.get('Rb', 5.0)
        ne = em['ne0'] * np.power(x, -em.get('alpha_n', 0.7))
        Te = em['Te0'] * np.power(x, -em.get('alpha_T', 1.0))
        B = em['B0'] * np.power(x, -em.get('alpha_B', 1.5))
        if em.get('bfield_model', False):
            B = np.asarray(Bmag, dtype=float)
        Theta_e = _KB * Te / (_ME * _C * _C)
        nu_c = 1.5 * (_E_CGS * B / (2.0 * np.pi * _ME * _C)) * Theta_e ** 2 * sinthb
        xx = nu_em / np.maximum(nu_c, 1e-40)
        II = 2.5651 * (1.0 + 1.92 * xx ** (-1.0 / 3.0) + 0.9977 * xx ** (-2.0 / 3.0)) * np.exp(-1.8899 * xx ** (1.0 / 3.0))
        return (ne * _E_CGS ** 2 * nu_em) / (2 * np.sqrt(3) * _C * (Theta_e ** 2)) * II
    if t == 'constant':
        j = np.ones(r.shape)
    elif t == 'bpl':
        rh = 1 + np.sqrt(1 - a ** 2)
        j = np.exp(em.get('p1', -2.0) * np.log(r / rh) + em.get('p2', -0.5) * np.log(r / rh) ** 2)
    elif t in ('glm', 'ring'):
        if t == 'ring':
            mu, goff, sig = em.get('r_ring', 4.5), em.get('gamma_off', 0), em.get('sigma', 0.3)
        else:
            mu, goff, sig = False, em.get('gamma_off', -1.5), em.get('sigma', 0.5)
        if not mu:
            mu = 1 - np.sqrt(1 - a ** 2)
        j = np.exp(-0.5 * (goff + np.arcsinh((r - mu) / sig)) ** 2) / np.sqrt((r - mu) ** 2 + sig ** 2)
    else:
        raise Exception("emistype %s not in oracle" % t)
    if g is not None:
        j = j * (g ** specind)
    if sinthb is not None:
        j = j * (sinthb ** (1 + specind))
    return j


# --------------------------------------------------------------------------------------
# equatorial image pipeline                          (kgeo/equatorial_images.py)
# --------------------------------------------------------------------------------------
def _photon_potentials(a, r, lam, eta, th=np.pi / 2):
    a2, r2 = a ** 2, r ** 2
    cth2, sth2 = np.cos(th) ** 2, np.sin(th) ** 2
    Delta = r2 - 2 * r + a2
    R = (r2 + a2 - a * lam) ** 2 - Delta * (eta + (lam - a) ** 2)
    TH = eta + a2 * cth2 - lam * lam * cth2 / sth2
    return Delta, R, TH


def redshift(a, r, lam, eta, kr_sign, kth_sign, u0, u1, u2, u3, th=np.pi / 2):
    """kgeo/equatorial_images.py:307-334."""
    Delta, R, TH = _photon_potentials(a, r, lam, eta, th)
    return 1 / (u0 - lam * u3 - np.sign(kth_sign) * u2 * np.sqrt(TH) - np.sign(kr_sign) * u1 * np.sqrt(R) / Delta)


def pol_quantities(a, r, lam, eta, kr_sign, kth_sign, vel, bf, th=np.pi / 2):
    """sin(theta_b), Penrose-Walker kappa, path length, b^2 (kgeo/equatorial_images.py:336-487)."""
    a2, r2 = a ** 2, r ** 2
    cth2, sth2 = np.cos(th) ** 2, np.sin(th) ** 2
    Delta, Sigma, g00, g11, g22, g33, g03 = _metric_eq(a, r, th)
    gdet = Sigma * np.sqrt(sth2)
    g00u = -(r2 + a2 + 2 * r * a2 * sth2 / Sigma) / Delta
    g11u = Delta / Sigma
    g22u = 1. / Sigma
    g33u = (Delta - a2 * sth2) / (Sigma * Delta * sth2)
    g03u = -(2 * a * r) / (Sigma * Delta)
    u0, u1, u2, u3 = velocity_u_lab(vel, a, r, th)
    l0, l1, l2, l3 = g00 * u0 + g03 * u3, g11 * u1, g22 * u2, g33 * u3 + g03 * u0
    (_, ex, ey, ez) = velocity_tetrads(vel, a, r, th)
    if bf['type'] == 'const_comoving':
        # the reference's comoving branch never defines norm_x (equatorial_images.py:422-436) and fails
        raise Exception("const_comoving is not usable through Iobs in the reference")
    B1, B2, B3 = bfield_lab(bf, a, r) if np.all(np.asarray(th) == np.pi / 2) else bfield_lab_omega(bf, a, r, th)[:3]
    n_l = (0, -B2 * gdet, B1 * gdet, 0)
    b0 = B1 * l1 + B2 * l2 + B3 * l3
    b1, b2, b3 = (B1 + b0 * u1) / u0, (B2 + b0 * u2) / u0, (B3 + b0 * u3) / u0
    bl = (g00 * b0 + g03 * b3, g11 * b1, g22 * b2, g33 * b3 + g03 * b0)

    def project(e, v):
        return e[0] * v[0] + e[1] * v[1] + e[2] * v[2] + e[3] * v[3]

    Bx, By, Bz = project(ex, bl), project(ey, bl), project(ez, bl)
    nx, ny, nz = project(ex, n_l), project(ey, n_l), project(ez, n_l)
    Bmag = np.sqrt(Bx ** 2 + By ** 2 + Bz ** 2)
    bsq = Bmag ** 2
    nmag = np.sqrt(nx ** 2 + ny ** 2 + nz ** 2)
    _, R, TH = _photon_potentials(a, r, lam, eta, th)
    kl = (-1, kr_sign * np.sqrt(R) / Delta, kth_sign * np.sqrt(TH), lam)
    k0 = g00u * kl[0] + g03u * kl[3]
    k1 = g11u * kl[1]
    k2 = g22u * kl[2]
    k3 = g33u * kl[3] + g03u * kl[0]
    kx, ky, kz = project(ex, kl), project(ey, kl), project(ez, kl)
    kmag = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    kdotn = kx * nx / nmag + ky * ny / nmag + kz * nz / nmag
    pathlength = np.abs(kmag / kdotn)
    fx = (ky * Bz - kz * By) / kmag
    fy = (kz * Bx - kx * Bz) / kmag
    fz = (kx * By - ky * Bx) / kmag
    sinthb = np.sqrt(fx ** 2 + fy ** 2 + fz ** 2) / Bmag
    f0 = ex[0] * fx + ey[0] * fy + ez[0] * fz
    f1 = ex[1] * fx + ey[1] * fy + ez[1] * fz
    f2 = ex[2] * fx + ey[2] * fy + ez[2] * fz
    f3 = ex[3] * fx + ey[3] * fy + ez[3] * fz
    A = k0 * f1 - k1 * f0 + a * sth2 * (k1 * f3 - k3 * f1)
    B = ((r2 + a2) * (k3 * f2 - k2 * f3) - a * (k0 * f2 - k2 * f0)) * np.sin(th)
    kappa = (r - 1j * a * np.cos(th)) * (A - 1j * B)
    return sinthb, kappa, pathlength, bsq


def evpa_terms(a, th_o, alpha, beta, kappa):
    """cos 2chi, sin 2chi by Penrose-Walker transport (kgeo/equatorial_images.py:524-544)."""
    mu = -(alpha + a * np.sin(th_o))
    k1, k2 = kappa.real, kappa.imag
    den = (beta ** 2 + mu ** 2) * (k1 ** 2 + k2 ** 2)
    c2 = ((beta * k1 + mu * k2) ** 2 - (mu * k1 - beta * k2) ** 2) / den
    s2 = (2 * (beta * k1 + mu * k2) * (mu * k1 - beta * k2)) / den
    return c2, s2


def iobs(a, r_o, th_o, mbar, alpha, beta, em, vel, bf, polarization=False, pathlength=False,
         specind=1, diskangle=0.1, nu_obs=230.e9):
    """Iobs (kgeo/equatorial_images.py:124-264) -> dict of per-pixel arrays."""
    alpha, beta = _arr(alpha), _arr(beta)
    rh = 1 + np.sqrt(1 - a ** 2)
    lam, eta = conserved(a, th_o, alpha, beta)
    r_s, Ir, Imax, Nmax = r_equatorial(a, r_o, th_o, mbar, alpha, beta)
    g = np.zeros(alpha.shape)
    sthb = np.zeros(alpha.shape)
    I = np.zeros(alpha.shape)
    Q = np.zeros(alpha.shape)
    U = np.zeros(alpha.shape)
    Ie = np.zeros(alpha.shape)
    lp = np.zeros(alpha.shape)
    dead = (Nmax < mbar) | (Nmax == -1) | (r_s <= rh)                      # :175
    cut_in, cut_out = em.get('emiscut_in', 0), em.get('emiscut_out', 1.e10)
    if cut_in > 0:
        dead = dead | (r_s < cut_in)
    if cut_out > 0:
        dead = dead | (r_s > cut_out)
    live = ~dead
    if live.any():
        rr, ll, ee = r_s[live], lam[live], eta[live]
        inside = ~(outside_crit(a, th_o, alpha[live], beta[live]).astype(bool))
        kr = np.where(inside | (Ir[live] <= 0.5 * Imax[live]), 1.0, -1.0)   # :266-288
        if th_o < 0.5 * np.pi:                                              # :290-305
            kth = -1 * np.power(-1, np.mod(mbar, 2))
        else:
            kth = 1 * np.power(-1, np.mod(mbar, 2))
        u0, u1, u2, u3 = velocity_u_lab(vel, a, rr)
        gg = redshift(a, rr, ll, ee, kr, kth, u0, u1, u2, u3)
        g[live] = gg
        if polarization:
            sb, kappa, _, bsq = pol_quantities(a, rr, ll, ee, kr, kth, vel, bf)
            c2, s2 = evpa_terms(a, th_o, alpha[live], beta[live], kappa)
        else:
            sb = np.pi / 4.                                                 # :20,:206
            bsq = bfield_bsq(bf, vel, a, rr)
        sthb[live] = sb
        Bmag = np.sqrt(bsq)
        if em['type'] == 'thermal':                                         # :221-223
            B0 = np.sqrt(bfield_bsq(bf, vel, a, _arr(em.get('Rb', 5.0))))
            Bmag = (Bmag / B0) * em['B0']
        je = emissivity_jrest(em, a, rr, gg, sb, Bmag=Bmag, nu_obs=nu_obs, specind=specind)
        if pathlength:                                                      # :489-522
            _, _, TH = _photon_potentials(a, rr, ll, ee)
            kyhat = -(kth * np.sqrt(TH)) / rr
            llp = (diskangle * rr) * ((1 / gg) / kyhat)
            je = je * np.abs(llp)
            I[live] = (gg ** 3) * je
            Ie[live] = je
            lp[live] = llp
        else:
            I[live] = (gg ** 2) * je
        if polarization:
            Q[live] = c2 * I[live]
            U[live] = s2 * I[live]
    return dict(I=I, Q=Q, U=U, g=g, r=r_s, sinthb=sthb, Ir=Ir, Imax=Imax, Nmax=Nmax, lp=lp, Ie=Ie)


def image_grid(alpha_min, alpha_max, beta_min, beta_max, psize):
    """Pixel grid of make_image, quirk included (kgeo/equatorial_images.py:48-55)."""
    n_alpha = int(np.floor(alpha_max - alpha_min) / psize)
    alphas = np.linspace(alpha_min, alpha_min + n_alpha * psize, n_alpha)
    n_beta = int(np.floor(beta_max - beta_min) / psize)
    betas = np.linspace(beta_min, beta_min + n_beta * psize, n_beta)
    A, B = np.meshgrid(alphas, betas)
    return A.flatten(), B.flatten()


def make_image(a, r_o, th_o, mbar_max, alpha_min, alpha_max, beta_min, beta_max, psize, em, vel, bf,
               polarization=False, pathlength=False, specind=1, diskangle=0.1, nu_obs=230.e9):
    """make_image (kgeo/equatorial_images.py:23-122) -> tuple in the reference's order."""
    A, B = image_grid(alpha_min, alpha_max, beta_min, beta_max, psize)
    shape = (len(A), mbar_max + 1)
    keys = ('I', 'Q', 'U', 'r', 'Ir', 'g', 'sinthb')
    planes = {k: np.zeros(shape) for k in keys + ('lp', 'Ie')}
    npol = nmax_poloidal(a, r_o, th_o, A, B)
    nmax = np.zeros(len(A))
    for mbar in range(mbar_max + 1):
        res = iobs(a, r_o, th_o, mbar, A, B, em, vel, bf, polarization, pathlength, specind, diskangle, nu_obs)
        for k in planes:
            planes[k][:, mbar] = res[k]
        nmax = res['Nmax']
    out = tuple(planes[k] for k in keys) + (nmax, npol)
    if pathlength:
        out = out + (planes['lp'], planes['Ie'])
    return out


# ---- field-line crossings (kgeo/solver.py) --------------------------------------------------------------------------
def psifunc(r, theta, a, model='para', pval=1):
    """kgeo/solver.py:13-37"""
    abscth = np.abs(np.cos(theta))
    if model == 'para':
        rp = 1 + np.sqrt(1 - a ** 2)
        return r * (1 - abscth) + rp * (1 + abscth) * (1 - np.log(1 + abscth)) - 2 * rp * (1 - np.log(2))
    if model == 'mono':
        return 1 - abscth
    if model == 'power':
        return r ** pval * (1 - abscth)
    return 0 * r


def maxtau_forwardjet(a, r_o, th_o, alpha, beta, neqmax=1, maxtaufrac=1.0 - 1.0e-10):
    """get_maxtau_forwardjet, kgeo/solver.py:57-88"""
    lam, eta = conserved(a, th_o, alpha, beta)
    r1, r2, r3, r4, _ = quartic_roots(a, lam, eta)
    tt2 = mino_total(a, r_o, eta, r1, r2, r3, r4) * maxtaufrac
    if neqmax is None:
        return tt2
    up, um, uratio, a2um = angular_consts(a, th_o, lam, eta)
    F0 = sp.ellipkinc(np.arcsin(np.cos(th_o) / np.sqrt(up)), uratio)
    K = sp.ellipkinc(0.5 * np.pi, uratio)
    mmax = neqmax - 1 + np.heaviside(beta, 0)
    taumax = (2 * mmax * K - np.sign(beta) * F0) / np.sqrt(-um * a ** 2)
    taumax = np.where(eta <= 0, tt2, taumax)
    return np.min(np.array([taumax, tt2]), axis=0)


def fieldline_guesses(r_s, th_s, tau_s, a, r_o, th_o, alpha, beta, psitarget, neqmax=1, model='para', pval=1):
    """getguesses, kgeo/solver.py:98-137, on (ngeo, npix) track arrays"""
    taumaxes = maxtau_forwardjet(a, r_o, th_o, alpha, beta, neqmax)
    psig = psifunc(r_s, th_s, a, model, pval) - psitarget
    out = []
    for i in range(len(alpha)):
        impact = np.sqrt(alpha[i] ** 2 + beta[i] ** 2)
        if np.abs(np.sin(th_o)) < .1 and impact > 100:
            if model == 'power' and pval > 0:
                rparam = (impact ** 2 / (2 * psitarget)) ** (1 / (2 - pval))
                out.append(np.array([1 / rparam, 2 / impact - 1 / rparam]))
                continue
            elif model == 'para':
                rparam = (impact ** 2 + psitarget ** 2) / (2 * psitarget)
                out.append(np.array([1 / rparam, 2 / impact - 1 / rparam]))
                continue
        minfunc = np.roll(psig[:, i], -1) * psig[:, i]
        indmax = np.argmin(np.abs(tau_s[:, i] - taumaxes[i]))
        minfunc = minfunc[:indmax + 1][1:-2]
        idx = np.where(minfunc < 0)[0] + 1
        if len(idx) == 0:
            out.append(np.array([-1]))
            continue
        out.append(np.array([(tau_s[k, i] + tau_s[k + 1, i]) / 2 for k in idx]))
    m = max(len(x) for x in out)
    res = -np.ones((len(out), m))
    for i, x in enumerate(out):
        res[i, :len(x)] = x
    return np.transpose(res.copy())


def _secant_all(func, x0, tol, maxiter=500):
    """scipy.optimize.newton(func, x0) for arrays without a derivative (scipy/optimize/_zeros_py.py::_array_newton, secant
    branch): every element keeps iterating until all have converged."""
    p = np.array(x0, copy=True)
    failures = np.ones_like(p, dtype=bool)
    dx = np.finfo(float).eps ** 0.33
    p1 = p * (1 + dx) + np.where(p >= 0, dx, -dx)
    q0, q1 = np.asarray(func(p)), np.asarray(func(p1))
    active = np.ones_like(p, dtype=bool)
    for _ in range(maxiter):
        nz = (q1 != q0)
        if not nz.any():
            p = (p1 + p) / 2.0
            break
        dp = (q1 * (p1 - p))[nz] / (q1 - q0)[nz]
        p[nz] = p1[nz] - dp
        azd = ~nz & active
        p[azd] = (p1 + p)[azd] / 2.0
        active &= nz
        failures[nz] = np.abs(dp) >= tol
        if not failures[nz].any():
            break
        p1, p = p, p1
        q0 = q1
        q1 = np.asarray(func(p1))
    return p


def findroot(r_s, th_s, tau_s, psitarget, alpha, beta, r_o, th_o, a, model='para', neqmax=1, tol=1e-8, pval=1):
    """findroot, kgeo/solver.py:142-207, on the (ngeo, npix) tracks of raytrace()"""
    guesses = fieldline_guesses(r_s, th_s, tau_s, a, r_o, th_o, alpha, beta, psitarget, neqmax, model, pval)
    shape = guesses.shape
    nrep = len(guesses)
    al, be = np.tile(alpha, nrep), np.tile(beta, nrep)
    g = guesses.flatten()
    obs = [0, r_o, th_o, 0]

    def coords(tau):
        _, (_, r, th, _) = coords_at_tau(a, obs, [al, be], tau.reshape(1, -1), full=False)
        return r[0], th[0]

    def func(tau):
        r, th = coords(tau)
        v = psifunc(r, th, a, model, pval) - psitarget
        v[g == -1] = 0
        return v

    root = _secant_all(func, g, tol)
    r0, t0 = coords(root)
    r1, t1 = coords(root * (1 + 1e-5))
    signpr, signpth = np.sign(r0 - r1), np.sign(t0 - t1)
    lam, eta = conserved(a, th_o, al, be)
    up, um, uratio, _ = angular_consts(a, th_o, lam, eta)
    F0 = sp.ellipkinc(np.arcsin(np.cos(th_o) / np.sqrt(up)), uratio)
    K = sp.ellipkinc(0.5 * np.pi, uratio)
    mbar = (np.sqrt(-um * a ** 2) * root + np.sign(be) * F0) / (2 * K) - np.heaviside(be, 0)
    neq = np.ceil(np.nan_to_num(mbar))
    none = (g == -1)
    root, r0, t0, signpr, signpth = [np.where(none, 0.0, x) for x in (root, r0, t0, signpr, signpth)]
    return root, r0, t0, signpr, signpth, neq, shape


# ---- jet-model densities (kgeo/densityfuncs.py, kgeo/ff_boost.py) -- groundwork for the rest of SURVEY 8 f4 ---------------
def r0min_mono(theta, Omegaf, spin, M=1.0):
    """stagnation surface of the monopole: largest-index root of the quintic of kgeo/ff_boost.py:8-19"""
    theta = np.arccos(np.abs(np.cos(theta)))
    c2t = np.cos(2 * theta)
    fac = (2 - spin * Omegaf + spin * Omegaf * c2t) ** 2
    coefs = [-spin ** 2 / 2 * M * np.cos(theta) ** 2 * fac, -2 * spin ** 4 * Omegaf ** 2 * np.cos(theta) ** 4 * np.sin(theta) ** 2,
             M / 2 * fac, -spin ** 2 * Omegaf ** 2 * np.sin(2 * theta) ** 2, 0.0, -2 * Omegaf ** 2 * np.sin(theta) ** 2]
    return np.real(np.polynomial.polynomial.Polynomial(coefs).roots()[4])


def eta_mono(rvals, theta, r0, spin, sigma):
    """mass loading of the monopole jet: Gaussian source integrated in closed form (kgeo/densityfuncs.py:31-41)"""
    if sigma == 0:
        return np.sign(rvals - r0)
    cth = np.cos(theta)
    s2pi = np.sqrt(2 * np.pi)
    num1 = 2 * sigma * (2 * r0 - (rvals + r0) * np.exp(-(rvals - r0) ** 2 / (2 * sigma ** 2)))
    num2 = s2pi * sp.erf((rvals - r0) / (np.sqrt(2) * sigma)) * (r0 ** 2 + sigma ** 2 + spin ** 2 * cth ** 2)
    return (num1 + num2) / (4 * r0 * sigma + s2pi * sigma ** 2 + s2pi * (r0 ** 2 + spin ** 2 * cth ** 2))


def density_mono_all(rvals, thvals, shape, spin, nu_parallel, sigma):
    """density_mono_all, kgeo/densityfuncs.py:44-79: rho = eta B^r / u^r per crossing, crossings reflected to the north"""
    bf = dict(type='bz_monopole')
    rnew = np.reshape(rvals, shape)
    thnew = np.arccos(np.abs(np.cos(np.reshape(thvals, shape))))
    i0 = np.where(rnew[0] != 0)[0][0]
    omega = bfield_lab_omega(bf, spin, rnew[0][i0:i0 + 1], thnew[0][i0])[3][0]
    r0 = r0min_mono(thnew[0][i0], omega, spin, 1.0)
    rho = []
    for ind in range(len(rnew)):
        B1 = bfield_lab_omega(bf, spin, rnew[ind], thnew[ind])[0]
        u1 = u_driftframe(spin, rnew[ind], bf, nu_parallel, thnew[ind])[1]
        rho.append(eta_mono(rnew[ind], thnew[ind], r0, spin, sigma) * B1 / u1)
    return np.nan_to_num(np.array(rho).flatten(), nan=0.0)


def _nderiv(r, theta, a, Omegaf, M, bf):
    """B . grad N along the field line (kgeo/ff_boost.py:85-99)"""
    cth, sth = np.cos(theta), np.sin(theta)
    Sigma = r ** 2 + a ** 2 * cth ** 2
    dNdr = (a ** 4 * Omegaf ** 2 * r * cth ** 4 * sth ** 2
            + r ** 2 * (-M + Omegaf * sth ** 2 * (2 * a * M + Omegaf * r ** 3 - a ** 2 * M * Omegaf * sth ** 2))
            + a ** 2 * cth ** 2 * (M + Omegaf * sth ** 2 * (-2 * a * M + 2 * Omegaf * r ** 3 + a ** 2 * M * Omegaf * sth ** 2)))
    dNdr /= Sigma ** 2 / 2
    den = a ** 2 + 2 * r ** 2 + a ** 2 * np.cos(2 * theta) ** 2
    dNdth = Omegaf ** 2 * (a ** 2 + r * (r - 2 * M)) + 8 * M * r * (a * (a * Omegaf - 1) + Omegaf * r ** 2) ** 2 / den ** 2
    dNdth *= np.sin(2 * theta)
    B = bfield_lab_omega(bf, a, np.array([r]).flatten(), theta)
    return (B[0] * dNdr + B[1] * dNdth)[0]


def _rfromR_power(R, psi0, p):
    """kgeo/ff_boost.py:74-83 (SciPy's scalar secant, as in the reference)"""
    import scipy.optimize as opt

    def f(Z):
        return np.sqrt(R ** 2 + Z ** 2) ** p * (1 - np.abs(np.cos(np.arctan(R / Z)))) - psi0
    try:
        Z = opt.newton(f, 4.0)
    except Exception:
        Z = opt.newton(f, 0)
    return np.sqrt(Z ** 2 + R ** 2)


def r0min_power(psi, Omegaf, spin, p, M=1.0, usemono=False):
    """stagnation surface of the power-law jet (kgeo/ff_boost.py:34-45)"""
    import scipy.optimize as opt
    bf = dict(type='power', p=p, usemono=usemono)

    def f(R):
        rs = _rfromR_power(R, psi, p)
        return _nderiv(rs, np.arcsin(R / rs), spin, Omegaf, M, bf)
    Rt = opt.newton(f, 4)
    rt = _rfromR_power(Rt, psi, p)
    return rt, np.arcsin(Rt / rt)


def eta_power(rvals, r0, spin, pval, psi, sigma):
    """mass loading of the power-law jet by quadrature of the Gaussian source (kgeo/densityfuncs.py:24-28,155-183)"""
    import scipy.integrate as scint
    if sigma == 0:
        return np.sign(rvals - r0)

    def src(r):
        return np.exp(-(r - r0) ** 2 / (2 * sigma ** 2)) * (r ** 2 + spin ** 2 * (1 - psi / r ** pval) ** 2)
    norm = scint.quad(src, r0, np.inf, epsabs=1e-13, epsrel=1e-13)[0]
    i0 = np.where(rvals != 1)[0][0]
    vals = [0.0] * i0
    for i in range(i0, len(rvals)):
        vals.append(0 if rvals[i] == 0 else scint.quad(src, r0, rvals[i], epsabs=1e-13, epsrel=1e-13)[0])
    return np.array(vals) / norm


def density_power_all(rvals, thvals, shape, psitarget, spin, nu_parallel, sigma, pval=1.0, usemono=False):
    """density_power_all, kgeo/densityfuncs.py:186-219"""
    bf = dict(type='power', p=pval, usemono=usemono)
    rnew = np.reshape(rvals, shape)
    thnew = np.arccos(np.abs(np.cos(np.reshape(thvals, shape))))
    i0 = np.where(np.nan_to_num(rnew)[0] != 0)[0][0]
    omega = np.broadcast_to(bfield_lab_omega(bf, spin, rnew[0][i0:i0 + 1], thnew[0][i0])[3], (1,))[0]
    r0, _ = r0min_power(psitarget, omega, spin, pval, 1.0, usemono)
    rho = []
    for ind in range(len(rnew)):
        B1 = bfield_lab_omega(bf, spin, rnew[ind], thnew[ind])[0]
        u1 = u_driftframe(spin, rnew[ind], bf, nu_parallel, thnew[ind])[1]
        rho.append(eta_power(rnew[ind], r0, spin, pval, psitarget, sigma) * B1 / u1)
    return np.nan_to_num(np.array(rho).flatten(), nan=0.0)


def iobs_off(a, r_o, r_s, th_o, alpha, beta, kr_sign, kth_sign, vel, bf, th_s, density=1, specind=1):
    """Iobs_off with Emissivity('constant') and polarization=True as getstokes calls it (kgeo/off_eq.py:15-121):
    (I, Q, U, sin theta_b) per crossing."""
    alpha, beta, r_s = _arr(alpha), _arr(beta), _arr(r_s)
    rh = 1 + np.sqrt(1 - a ** 2)
    lam, eta = conserved(a, th_o, alpha, beta)
    u0, u1, u2, u3 = velocity_u_lab(vel, a, r_s, th_s)
    gg = redshift(a, r_s, lam, eta, kr_sign, kth_sign, u0, u1, u2, u3, th_s)
    Iemis = emissivity_jrest(dict(type='constant'), a, r_s) * density
    sinthb, kappa, pathlength, bsq = pol_quantities(a, r_s, lam, eta, kr_sign, kth_sign, vel, bf, th_s)
    c2, s2 = evpa_terms(a, th_o, alpha, beta, kappa)
    Iemis = Iemis * bsq ** ((1 + specind) / 2)
    I = (gg ** 3) * (gg ** specind) * pathlength * Iemis * (sinthb ** (1 + specind))
    Q, U = c2 * I, s2 * I
    inside = r_s <= rh
    I, Q, U = [np.where(inside, 0, x) for x in (I, Q, U)]
    return I, Q, U, sinthb


def sort_image(iobs, qobs, uobs, neqvals, shape, ashape, neqmax):
    """geometry.sort_image (kgeo/geometry.py:38-62): one image per crossing order, the total last"""
    out = []
    for arr in (iobs, qobs, uobs):
        planes = [np.reshape(np.sum(np.reshape(np.where(neqvals != neq, 0, arr), shape), axis=0), ashape) for neq in range(neqmax)]
        planes.append(np.reshape(np.sum(np.reshape(arr, shape), axis=0), ashape))
        out.append(np.array(planes))
    return out


def getstokes(psitarget, alphavals, betavals, r_o, th_o, a, ngeo, neqmax=1, tol=1e-8, model='mono', pval=1, nu_parallel=0, sigma=2,
              sumsubring=True, gammamax=None):
    """off_eq.getstokes for the monopole and power-law jets (kgeo/off_eq.py:124-233); nu_parallel numeric or (monopole) 'FF'"""
    ashape = alphavals.shape
    A, B = alphavals.flatten(), betavals.flatten()
    tau_s, _, (_, r_t, th_t, _) = raytrace(a, [0, r_o, th_o, 0], [A, B], ngeo)
    tau, rvals, thvals, signpr, signpth, neqvals, shape = findroot(r_t, th_t, tau_s, psitarget, A, B, r_o, th_o, a, model, neqmax,
                                                                   tol, pval)
    At, Bt = np.tile(A, shape[0]), np.tile(B, shape[0])
    if model == 'mono' or (model == 'power' and pval == 0):
        bf = dict(type='bz_monopole')
        dvals = density_mono_all(rvals, thvals, shape, a, nu_parallel, sigma)
    else:
        bf = dict(type='power', p=pval)
        dvals = density_power_all(rvals, thvals, shape, psitarget, a, nu_parallel, sigma, pval)
    vel = dict(type='driftframe', bfield=bf, nu_parallel=nu_parallel, gammamax=gammamax)
    I, Q, U, sinthb = iobs_off(a, r_o, rvals, th_o, At, Bt, signpr, signpth, vel, bf, thvals, dvals, 1)
    none = (rvals == 0)
    I, Q, U = [np.real(np.nan_to_num(np.where(none, 0, x))) for x in (I, Q, U)]
    if not sumsubring:
        return I, Q, U, neqvals, np.nan_to_num(rvals), np.nan_to_num(thvals), shape
    iv, qv, uv = sort_image(I, Q, U, neqvals, shape, ashape, neqmax)
    return iv, qv, uv, np.nan_to_num(0.5 * np.arctan2(uv, qv)), np.nan_to_num(rvals), np.nan_to_num(thvals), np.nan_to_num(sinthb)
