"""Drop-in for kgeo/off_eq.py getstokes: polarized images of off-equatorial emission from one field line of a BZ jet model.

Per-crossing work (field-line crossings, densities, drift-frame velocity, redshift, Penrose-Walker transport, Stokes parameters,
subring sums) runs in libkgeo_b200.so kernels; what stays on the host is one scalar chain per call, as in the reference: the
field-line angular speed at the first crossing, the stagnation surface (kgeo/ff_boost.py: a quintic for the monopole, two nested
secant solves for the power law), the normalisation of the power law's mass loading and, for nu_parallel='FF', the conserved
E - Omega_F L of the field line (ff_boost.getEco).

Parity with the unmodified reference's getstokes (monopole and power-law models): tests/test_gpu_jet.py.  The paraboloid model
raises: in the reference itself densityfuncs.eta_para fails on the complex numbers SciPy's lambertw produces.

getstokes -- kgeo/off_eq.py:124-233        density_*_all -- kgeo/densityfuncs.py:44-79,186-219        r0min_* -- kgeo/ff_boost.py:8-45
"""
import ctypes

import numpy as np

from . import _engine, _lib, solver
from .bfields import bz_fr_table
from .kerr_raytracing_ana import raytrace_ana

SPECIND = 1      # kgeo/off_eq.py:13
JET_MONO, JET_POWER = 0, 1


# ---- host scalars ---------------------------------------------------------------------------------------------------------------
def _secant(f, x0, tol=1.48e-8, maxiter=50):
    """scipy.optimize.newton(f, x0) without a derivative (scalar secant branch), statement for statement"""
    p0 = 1.0 * x0
    eps = 1e-4
    p1 = x0 * (1 + eps)
    p1 += (eps if p1 >= 0 else -eps)
    q0, q1 = f(p0), f(p1)
    if abs(q1) < abs(q0):
        p0, p1, q0, q1 = p1, p0, q1, q0
    for itr in range(maxiter):
        if q1 == q0:
            if p1 != p0:
                raise RuntimeError("Tolerance of %s reached. Failed to converge after %d iterations, value is %s." % (p1 - p0, itr + 1, p1))
            return (p1 + p0) / 2.0
        if abs(q1) > abs(q0):
            p = (-q0 / q1 * p1 + p0) / (1 - q0 / q1)
        else:
            p = (-q1 / q0 * p0 + p1) / (1 - q1 / q0)
        if np.isclose(p, p1, rtol=0.0, atol=tol):
            return p
        p0, q0 = p1, q1
        p1 = p
        q1 = f(p1)
    raise RuntimeError("Failed to converge after %d iterations, value is %s." % (maxiter, p))


def _omega_power(r, th, a, pval, usemono):
    """omega_BZpower (kgeo/bfields.py:618-625) on psi = r^p (1 - |cos th|)"""
    if usemono or pval == 0:
        return a / 8
    rp = 1 + np.sqrt(1 - a ** 2)
    psi = r ** pval * (1 - np.abs(np.cos(th)))
    return a / (4 + 8 / (1 + (1 - psi * rp ** (-pval))))


def _omega_mono(th, a):
    """field-line speed of the perturbative monopole (kgeo/bfields.py:567-590)"""
    f2 = (6 * np.pi * np.pi - 49.) / 72.
    omega2 = 1. / 32. - np.sin(th) ** 2 * (4 * f2 - 1) / 64.
    spin = np.abs(a)
    om = spin / 8. + spin ** 3 * omega2
    return -om if a < 0 else om


def r0min_mono(theta, Omegaf, spin, M=1.0):
    """stagnation radius of the monopole jet: root 4 of the quintic of kgeo/ff_boost.py:8-19"""
    theta = np.arccos(np.abs(np.cos(theta)))
    fac = (2 - spin * Omegaf + spin * Omegaf * np.cos(2 * theta)) ** 2
    coefs = [-spin ** 2 / 2 * M * np.cos(theta) ** 2 * fac, -2 * spin ** 4 * Omegaf ** 2 * np.cos(theta) ** 4 * np.sin(theta) ** 2,
             M / 2 * fac, -spin ** 2 * Omegaf ** 2 * np.sin(2 * theta) ** 2, 0.0, -2 * Omegaf ** 2 * np.sin(theta) ** 2]
    return float(np.real(np.polynomial.polynomial.Polynomial(coefs).roots()[4]))


def getEco(r0, theta, Omegaf, spin, M=1.0):
    """E - Omega_F L of the cold flow launched at rest from (r0, theta) (kgeo/ff_boost.py:102-119, metric of :122-146)"""
    Sig = r0 ** 2 + spin ** 2 * np.cos(theta) ** 2
    Del = r0 ** 2 - 2 * M * r0 + spin ** 2
    s2 = np.sin(theta) ** 2
    gtt = -(1 - 2 * M * r0 / Sig)
    gtphi = -2 * spin * M * r0 / Sig * s2
    gphiphi = (r0 ** 2 + spin ** 2 + 2 * M * r0 * spin ** 2 / Sig * s2) * s2
    ginvtt = -(r0 ** 2 + spin ** 2 + 2 * M * r0 * spin ** 2 / Sig * s2) / Del
    ginvtphi = -2 * M * r0 * spin / Sig / Del
    ginvphiphi = (Del - spin ** 2 * s2) / (Sig * Del * s2)
    gtphifac = gtphi + gphiphi * Omegaf
    gttfac = gtt + gtphi * Omegaf
    coef0 = (gtt + Omegaf * (2 * gtphi + gphiphi * Omegaf)) ** 2
    coef1 = ginvphiphi * gtphifac ** 2 + 2 * ginvtphi * gtphifac * gttfac + ginvtt * gttfac ** 2
    return float(np.sqrt(-coef0 / coef1))


def ff_constants_mono(r_first, th_first, spin):
    """(r0, omega, E - Omega_F L) of the monopole field line through (r_first, th_first): what u_driftframe(nu_parallel='FF')
    derives from the first crossing it is given (kgeo/velocities.py:495-516)"""
    omega = _omega_mono(th_first, spin)
    r0 = r0min_mono(th_first, omega, spin, 1.0)
    return r0, float(omega), getEco(r0, th_first, omega, spin, 1.0)


def _power_field_r_th(a, r, th, pval, usemono):
    """B^r, B^theta of Bfield('power', C=1) at one point (kgeo/bfields.py:676-709)"""
    sth, cth = np.sin(th), np.cos(th)
    gdet = sth * (r ** 2 + a ** 2 * cth ** 2)
    psi = r ** pval * (1 - np.abs(cth))
    return np.sign(cth) * sth * r ** pval / gdet, -(pval * psi / r) / gdet


def _rfromR_power(R, psi0, p):
    """kgeo/ff_boost.py:74-83"""
    def f(Z):
        return np.sqrt(R ** 2 + Z ** 2) ** p * (1 - np.abs(np.cos(np.arctan(R / Z)))) - psi0
    try:
        Z = _secant(f, 4.0)
    except Exception:
        Z = _secant(f, 0)
    return np.sqrt(Z ** 2 + R ** 2)


def r0min_power(psi, Omegaf, spin, p, M=1.0, usemono=False):
    """stagnation surface of the power-law jet: zero of B . grad N along the field line (kgeo/ff_boost.py:34-45,85-99)"""
    def nderiv(R):
        r = _rfromR_power(R, psi, p)
        theta = np.arcsin(R / r)
        cth, sth = np.cos(theta), np.sin(theta)
        a = spin
        Sigma = r ** 2 + a ** 2 * cth ** 2
        dNdr = (a ** 4 * Omegaf ** 2 * r * cth ** 4 * sth ** 2
                + r ** 2 * (-M + Omegaf * sth ** 2 * (2 * a * M + Omegaf * r ** 3 - a ** 2 * M * Omegaf * sth ** 2))
                + a ** 2 * cth ** 2 * (M + Omegaf * sth ** 2 * (-2 * a * M + 2 * Omegaf * r ** 3 + a ** 2 * M * Omegaf * sth ** 2)))
        dNdr /= Sigma ** 2 / 2
        den = a ** 2 + 2 * r ** 2 + a ** 2 * np.cos(2 * theta) ** 2
        dNdth = Omegaf ** 2 * (a ** 2 + r * (r - 2 * M)) + 8 * M * r * (a * (a * Omegaf - 1) + Omegaf * r ** 2) ** 2 / den ** 2
        dNdth *= np.sin(2 * theta)
        B1, B2 = _power_field_r_th(a, r, theta, p, usemono)
        return B1 * dNdr + B2 * dNdth
    Rt = _secant(nderiv, 4)
    rt = _rfromR_power(Rt, psi, p)
    return float(rt), float(np.arcsin(Rt / rt))


_GL_X = np.array([0.0950125098376374402, 0.2816035507792589132, 0.4580167776572273863, 0.6178762444026437484,
                  0.7554044083550030339, 0.8656312023878317439, 0.9445750230732325761, 0.9894009349916499326])
_GL_W = np.array([0.1894506104550684963, 0.1826034150449235888, 0.1691565193950025382, 0.1495959888165767321,
                  0.1246289712555338720, 0.0951585116824927848, 0.0622535239386478929, 0.0271524594117540949])


def power_source_norm(r0, spin, pval, psi, sigma):
    """integral of the mass-loading source from the stagnation radius to infinity (the `normfac` of kgeo/densityfuncs.py:167): the
    same 16-point Gauss-Legendre panels of width sigma, out to 40 sigma, that the device uses for the partial integrals"""
    total = 0.0
    for k in range(40):
        c, h = (k + 0.5) * sigma, 0.5 * sigma
        t = np.concatenate((c - h * _GL_X, c + h * _GL_X))
        x = r0 + t
        total += h * np.sum(np.concatenate((_GL_W, _GL_W)) * np.exp(-t * t / (2 * sigma ** 2)) * (x ** 2 + spin ** 2 * (1 - psi / x ** pval) ** 2))
    return float(total)


def jet_scalars(model, rvals, thvals, shape, psitarget, spin, sigma, pval=1.0, usemono=False):
    """(field id, r0, norm) of a getstokes call from the crossings of findroot: the reference takes the field-line speed at the
    first ray with a crossing (kgeo/densityfuncs.py:52-57,195-200)"""
    rnew = np.reshape(rvals, shape)
    thnew = np.arccos(np.abs(np.cos(np.reshape(thvals, shape))))
    i0 = np.where(np.nan_to_num(rnew)[0] != 0)[0][0]
    if model == 'mono' or (model == 'power' and pval == 0):
        return JET_MONO, r0min_mono(thnew[0][i0], _omega_mono(thnew[0][i0], spin), spin, 1.0), 1.0
    omega = _omega_power(rnew[0][i0], thnew[0][i0], spin, pval, usemono)
    r0, _ = r0min_power(psitarget, omega, spin, pval, 1.0, usemono=usemono)
    norm = power_source_norm(r0, spin, pval, psitarget, sigma) if sigma != 0 else 1.0
    return JET_POWER, r0, norm


# ---- the pipeline -------------------------------------------------------------------------------------------------------------------
def getstokes(psitarget, alphavals, betavals, r_o, th_o, a, ngeo, do_phi_and_t=True, neqmax=1, outgeo=None, tol=1e-8, model='para',
              pval=1, nu_parallel=0, gammamax=None, retvals=False, sigma=2, sumsubring=True, usemono=False, retsin=False):
    """Stokes parameters of the emission from the field line psi = psitarget of a BZ jet model, images ordered by the number of
    equatorial crossings with the total last (same arguments and return tuples as the reference)"""
    if model == 'para':
        raise NotImplementedError("getstokes(model='para'): the reference's own pipeline fails for the paraboloid "
                                  "(densityfuncs.eta_para on SciPy's complex lambertw values)")
    if model not in ('mono', 'power'):
        raise Exception("model %s not recognized in getstokes!" % model)
    nu_ff = isinstance(nu_parallel, str)
    if nu_ff and nu_parallel != 'FF':
        raise Exception("nu_parallel should be a number in [-1,1] or 'FF'")
    if nu_ff and not (model == 'mono' or (model == 'power' and pval == 0)):
        raise NotImplementedError("getstokes(nu_parallel='FF') is built for the monopole field only (the reference's own "
                                  "power-law path raises inside u_driftframe)")
    alphavals, betavals = np.asarray(alphavals, dtype=np.float64), np.asarray(betavals, dtype=np.float64)
    ashape = alphavals.shape
    alpha, beta = alphavals.flatten(), betavals.flatten()
    if outgeo is None:
        outgeo = raytrace_ana(a=a, observer_coords=[0, r_o, th_o, 0], image_coords=[alpha, beta], ngeo=ngeo,
                              do_phi_and_t=do_phi_and_t, savedata=False, plotdata=False)
    tau, rvals, thvals, signpr, signptheta, neqvals, shape = solver.findroot(outgeo, psitarget, alpha, beta, r_o, th_o, a, ngeo,
                                                                             model=model, neqmax=neqmax, tol=tol, pval=pval)
    print('guesses before ', shape)
    field, r0, norm = jet_scalars(model, rvals, thvals, shape, psitarget, a, sigma, pval, usemono)
    L = _lib.lib()
    dev = _engine.device()
    npix, nguess = int(shape[1]), int(shape[0])
    n = npix * nguess
    d = {k: _engine.to_device(np.ascontiguousarray(v, dtype=np.float64), dev)
         for k, v in dict(alpha=alpha, beta=beta, r=rvals, th=thvals, kr=signpr, kth=signptheta, neq=neqvals).items()}
    table = None
    if field == JET_MONO:
        rs, fs = bz_fr_table()
        table = _engine.to_device(np.ascontiguousarray(np.vstack((rs, fs))), dev)
    I, Q, U, sinthb = [_engine.empty((n,), dev) for _ in range(4)]
    obs = _engine.observer_struct(a, r_o, th_o)
    # nu_parallel = 'FF': the field-line constants of the two places the reference evaluates u_driftframe (host scalars, as the
    # reference derives them from the first crossing it is handed: kgeo/velocities.py:495-516)
    ff_global, ff_rows = None, None
    if nu_ff:
        rflat, thflat = np.nan_to_num(np.asarray(rvals, dtype=np.float64)), np.asarray(thvals, dtype=np.float64)
        i0 = np.where(rflat != 0)[0][0]                      # Iobs_off: flattened crossings, angles as they are
        ff_global = np.array(ff_constants_mono(rflat[i0], thflat[i0], a), dtype=np.float64)
        rnew = rflat.reshape(shape)
        thnew = np.arccos(np.abs(np.cos(thflat.reshape(shape))))
        rows = []
        for m in range(nguess):                              # density_mono_all: one u_driftframe call per crossing order
            im = np.where(rnew[m] != 0)[0][0]                # (IndexError when an order has no crossing at all, as there)
            rows.append(ff_constants_mono(rnew[m][im], thnew[m][im], a))
        ff_rows = _engine.to_device(np.ascontiguousarray(np.array(rows, dtype=np.float64).ravel()), dev)
    bad_nu = ctypes.c_int32(0)
    _lib.check(L.kgeo_jet_stokes(ctypes.byref(obs), npix, nguess, _engine.dptr(d['alpha']), _engine.dptr(d['beta']), _engine.dptr(d['r']),
                                 _engine.dptr(d['th']), _engine.dptr(d['kr']), _engine.dptr(d['kth']), field, 1.0, float(pval),
                                 int(bool(usemono)), _engine.dptr(table), 0 if table is None else int(table.shape[1]),
                                 0.0 if nu_ff else float(nu_parallel), int(nu_ff), _engine.hptr(ff_global), _engine.dptr(ff_rows),
                                 float(gammamax) if gammamax else 0.0, float(sigma), float(r0), float(norm), float(psitarget),
                                 float(SPECIND), _engine.dptr(I), _engine.dptr(Q), _engine.dptr(U), _engine.dptr(sinthb), None,
                                 ctypes.byref(bad_nu), _engine.stream_ptr()))
    if bad_nu.value:
        raise Exception("nu_parallel should be in the range (-1,1)")
    if not sumsubring:
        return (I.cpu().numpy(), Q.cpu().numpy(), U.cpu().numpy(), neqvals, np.nan_to_num(rvals), np.nan_to_num(thvals), shape)
    outs = [_engine.empty((neqmax + 1, npix), dev) for _ in range(3)]
    _lib.check(L.kgeo_jet_sort(npix, nguess, int(neqmax), _engine.dptr(I), _engine.dptr(Q), _engine.dptr(U), _engine.dptr(d['neq']),
                               *[_engine.dptr(o) for o in outs], _engine.stream_ptr()))
    ivec, qvec, uvec = [o.cpu().numpy().reshape((neqmax + 1,) + ashape) for o in outs]
    evpa = np.nan_to_num(0.5 * np.arctan2(uvec, qvec))
    if not retvals:
        return ivec, qvec, uvec, evpa
    if not retsin:
        return ivec, qvec, uvec, evpa, np.nan_to_num(rvals), np.nan_to_num(thvals)
    return ivec, qvec, uvec, evpa, np.nan_to_num(rvals), np.nan_to_num(thvals), sinthb.cpu().numpy()
