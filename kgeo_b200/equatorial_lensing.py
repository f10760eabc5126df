"""Drop-in for kgeo/equatorial_lensing.py: r_equatorial (:85-196), nmax_equatorial (:22-51), nmax_poloidal (:53-82) and the
inverse lensing of SURVEY 8(f2) -- rho_of_req (:226-247), rho_of_req_single (:249-266), critical_curve (:268-298) --
computed by libkgeo_b200.so (the root find runs on the GPU, one thread per image-plane angle).

generate_library / make_curves2 (:351-417, the reference's multiprocessing job over 99 x 89 configurations) run as ONE kernel
launch over (configuration, order, angle) -- rho_library -- and write the reference's text files.  Out of scope: plot_curves
(matplotlib).
"""
import os

import ctypes

import numpy as np

from . import _checks, _engine, _lib
from ._engine import dptr

RMAX = np.inf  # kgeo/equatorial_lensing.py:19
INF = 1.e50    # :18
R0 = np.inf    # :19
# scipy.optimize.brentq defaults, which the reference relies on (:243, :264)
BRENTQ_XTOL, BRENTQ_RTOL, BRENTQ_MAXITER = 2e-12, 4 * np.finfo(float).eps, 100


def _r_equatorial_device(a, r_o, th_o, mbar_lo, mbar_hi, alpha, beta, want=("rs", "Ir", "Nmax", "Imax", "npol")):
    L = _lib.lib()
    _engine.require_cuda()
    dev = _engine.device()
    obs = _engine.observer_struct(a, r_o, th_o)
    al, be = _engine.to_device(alpha, dev), _engine.to_device(beta, dev)
    n = int(al.numel())
    norder = mbar_hi - mbar_lo + 1
    out = {}
    for k in ("rs", "Ir", "Nmax"):
        out[k] = _engine.empty((norder, n), dev) if k in want else None
    for k in ("Imax", "npol"):
        out[k] = _engine.empty((n,), dev) if k in want else None
    _lib.check(L.kgeo_r_equatorial(ctypes.byref(obs), n, dptr(al), dptr(be), mbar_lo, mbar_hi, dptr(out["rs"]),
                                   dptr(out["Ir"]), dptr(out["Nmax"]), dptr(out["Imax"]), dptr(out["npol"]),
                                   _engine.stream_ptr()))
    return out


def nmax_equatorial(a, r_o, th_o, alpha, beta):
    """Return the maximum number of equatorial crossings"""
    _checks.check_observer(a, r_o, th_o)
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    out = _r_equatorial_device(a, r_o, th_o, 0, 0, alpha, beta, want=("Nmax",))
    return out["Nmax"][0].cpu().numpy()


def nmax_poloidal(a, r_o, th_o, alpha, beta):
    """Return the maximum number of poloidal orbits"""
    _checks.check_observer(a, r_o, th_o)
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    out = _r_equatorial_device(a, r_o, th_o, 0, 0, alpha, beta, want=("npol",))
    return out["npol"].cpu().numpy()


def r_equatorial(a, r_o, th_o, mbar, alpha, beta):
    """Return (r_s, Ir, Imax, Nmax) where

       r_s is the equatorial emission radius
       Ir is the elapsed Mino time at emission
       Imax is the maximal Mino time on the geodesic
       Nmax is the maximum number of equatorial crossings"""
    _checks.check_observer(a, r_o, th_o)
    if not (isinstance(mbar, int) and (mbar >= 0)):
        raise Exception("mbar should be an integer >=0!")
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    out = _r_equatorial_device(a, r_o, th_o, mbar, mbar, alpha, beta, want=("rs", "Ir", "Nmax", "Imax"))
    return (out["rs"][0].cpu().numpy(), out["Ir"][0].cpu().numpy(), out["Imax"].cpu().numpy(),
            out["Nmax"][0].cpu().numpy())


def r_equatorial_orders(a, r_o, th_o, mbar_max, alpha, beta):
    """Extension: all orders 0..mbar_max in one fused pass.  Returns (r_s, Ir, Nmax) of shape (mbar_max+1, npix) and
    Imax of shape (npix,)."""
    _checks.check_observer(a, r_o, th_o)
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    out = _r_equatorial_device(a, r_o, th_o, 0, int(mbar_max), alpha, beta, want=("rs", "Ir", "Nmax", "Imax"))
    return tuple(out[k].cpu().numpy() for k in ("rs", "Ir", "Nmax", "Imax"))


def _rho_bracket(req, mbar):
    # bounding ranges eyeballed from Gralla+Lupsasca Fig 6 (kgeo/equatorial_lensing.py:232-241)
    if mbar == 0:
        return 0, req + 5
    elif mbar == 1:
        return 0, 100
    return 1, 10


def _rho_of_req_device(a, th0, req, mbar, varphis):
    L = _lib.lib()
    _engine.require_cuda()
    t = _engine.torch()
    dev = _engine.device()
    obs = _engine.observer_struct(a, np.inf, th0)
    vp = _engine.to_device(np.atleast_1d(np.asarray(varphis, dtype=np.float64)), dev)
    n = int(vp.numel())
    rho, al, be = (_engine.empty((n,), dev) for _ in range(3))
    status = _engine.empty((n,), dev, dtype=t.int32)
    rhomin, rhomax = _rho_bracket(req, mbar)
    _lib.check(L.kgeo_rho_of_req(ctypes.byref(obs), n, dptr(vp), float(req), int(mbar), float(rhomin), float(rhomax),
                                 BRENTQ_XTOL, BRENTQ_RTOL, BRENTQ_MAXITER, dptr(rho), dptr(al), dptr(be),
                                 _engine.iptr(status), _engine.stream_ptr()))
    st = status.cpu().numpy()
    if (st == -1).any():      # what scipy.optimize.brentq raises inside the reference's list comprehension
        raise ValueError("f(a) and f(b) must have different signs")
    if (st == -2).any():
        raise RuntimeError("Failed to converge after %d iterations." % BRENTQ_MAXITER)
    return rho.cpu().numpy(), al.cpu().numpy(), be.cpu().numpy()


def rho_of_req(a, th0, req, mbar=0, varphis=None):
    """Image-plane curve (varphis, rhos, alphas, betas) of the mbar-th image of the equatorial ring r = req
    (kgeo/equatorial_lensing.py:226-247)"""
    if varphis is None:
        varphis = np.linspace(-180, 179, 360) * np.pi / 180.
    rhos, alphas, betas = _rho_of_req_device(a, th0, req, mbar, varphis)
    return (varphis, rhos, alphas, betas)


def rho_of_req_single(a, th0, req, varphi, mbar=0):
    """rho of one image-plane angle (kgeo/equatorial_lensing.py:249-266)"""
    rhos, _, _ = _rho_of_req_device(a, th0, req, mbar, [varphi])
    return float(rhos[0])


def r_equatorial2(rho, varphi, a, th0, mbar=0):
    """Emission radius of the mbar-th image at image-plane polar coordinates (rho, varphi), observer at infinity, made
    sign-definite for the root find: 1e50 where the ray misses the plane outside the critical curve, 1 inside
    (kgeo/equatorial_lensing.py:198-217).  Scalar in, scalar out, as the reference's."""
    from .kerr_raytracing_utils import is_outside_crit
    alpha = np.atleast_1d(np.cos(varphi) * rho).astype(float)
    beta = np.atleast_1d(np.sin(varphi) * rho).astype(float)
    r = r_equatorial(a, R0, th0, mbar, alpha, beta)[0][0]
    if r <= 0:
        r = INF if is_outside_crit(a, th0, alpha, beta)[0] == 1 else 1
    return r


def objfunc(rho, varphi, a, th0, r_target, mbar=0):
    """r_equatorial2 - r_target: the objective of the reference's Brent solve (:219-224)"""
    return r_equatorial2(rho, varphi, a, th0, mbar=mbar) - r_target


def rho_library(spins, th0s, reqs, mmax=5, varphis=None):
    """rho_of_req for MANY configurations in one kernel launch: spins[c], th0s[c] (radians), reqs[c] for c < nconf, orders
    mbar = 0..mmax, angles `varphis` (default: rho_of_req's 360).  Returns (varphis, rhos[nconf, mmax+1, nphi],
    status[nconf, mmax+1, nphi]) with status 0 = converged, -1 = no sign change over the bracket (SciPy's ValueError),
    -2 = not converged (SciPy's RuntimeError)."""
    L = _lib.lib()
    _engine.require_cuda()
    t = _engine.torch()
    dev = _engine.device()
    spins, th0s, reqs = [np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (spins, th0s, reqs)]
    if not (len(spins) == len(th0s) == len(reqs)):
        raise Exception("spins, th0s, reqs are different lengths!")
    for a, th0 in zip(spins, th0s):
        _checks.check_observer(float(a), np.inf, float(th0))
    if varphis is None:
        varphis = np.linspace(-180, 179, 360) * np.pi / 180.
    varphis = np.atleast_1d(np.asarray(varphis, dtype=np.float64))
    nconf, nphi, norder = len(spins), len(varphis), int(mmax) + 1
    d = [_engine.to_device(x, dev) for x in (spins, th0s, reqs, varphis)]
    rho = _engine.empty((nconf, norder, nphi), dev)
    status = _engine.empty((nconf, norder, nphi), dev, dtype=t.int32)
    _lib.check(L.kgeo_rho_library(nconf, dptr(d[0]), dptr(d[1]), dptr(d[2]), 0, int(mmax), nphi, dptr(d[3]), BRENTQ_XTOL,
                                  BRENTQ_RTOL, BRENTQ_MAXITER, dptr(rho), _engine.iptr(status), _engine.stream_ptr()))
    return varphis, rho.cpu().numpy(), status.cpu().numpy()


_LIBRARY_RADII = {'r10': 10., 'r100': 100., 'r1000': 1000., 'r10000': 10000.}


def _library_radius(which, spin):
    if which == 'rh':
        return 1 + np.sqrt(1 - spin ** 2)
    if which not in _LIBRARY_RADII:
        raise Exception()
    return _LIBRARY_RADII[which]


def _write_curves(outpath, which, spins, incs, varphis, rhos, status):
    written = 0
    for c, (spin, inc) in enumerate(zip(spins, incs)):
        name = '%s/a%02.0f_i%02.0f_%s.txt' % (outpath, 100 * spin, inc, which)
        if os.path.isfile(name) or (status[c] != 0).any():   # the reference skips existing files and drops failed configurations
            continue
        np.savetxt(name, np.vstack((varphis, rhos[c])).T)
        written += 1
    return written


def make_curves2(i, which, n, spin, inc, mmax, outpath):
    """One configuration of the library (kgeo/equatorial_lensing.py:379-417): the file '<outpath>/aXX_iYY_<which>.txt' with
    columns varphi, rho(mbar = 0), ..., rho(mbar = mmax); nothing is written when a root find fails or the file exists."""
    rs = _library_radius(which, spin)
    varphis, rhos, status = rho_library([spin], [inc * np.pi / 180.], [rs], mmax)
    _write_curves(outpath, which, [spin], [inc], varphis, rhos, status)


def make_curves(args):
    return make_curves2(*args)


def generate_library(which='rh', outpath=None, spins=None, incs=None, mmax=5):
    """The reference's curve library (:351-374): spins linspace(0.01, 0.99, 99) x inclinations 1..89 deg, orders 0..5, 360 angles,
    one text file per configuration under './curve_library_<which>'.  The 8811 x 6 x 360 root finds run as one kernel launch
    (the reference: a multiprocessing pool); returns the number of files written."""
    outpath = outpath or './curve_library_' + which
    print(outpath)
    spins = np.linspace(0.01, 0.99, 99) if spins is None else np.asarray(spins, dtype=float)
    incs = np.linspace(1, 89, 89) if incs is None else np.asarray(incs, dtype=float)
    allincs, allspins = np.meshgrid(incs, spins)
    allincs, allspins = allincs.flatten(), allspins.flatten()
    reqs = np.array([_library_radius(which, s) for s in allspins])
    os.makedirs(outpath, exist_ok=True)
    varphis, rhos, status = rho_library(allspins, allincs * np.pi / 180., reqs, mmax)
    return _write_curves(outpath, which, allspins, allincs, varphis, rhos, status)


def critical_curve(a, th0, n=100000):
    """returns parametrized critical curve (alpha,beta) array with n points
       won't work for a=0 exactly"""
    L = _lib.lib()
    _engine.require_cuda()
    t = _engine.torch()
    dev = _engine.device()
    # beginning and ending points, Eq 40 (scalars, exactly as the reference forms them: :275-276, :288-289)
    rminus = 2 * (1 + np.cos(2. * np.arccos(-a) / 3.))
    rplus = 2 * (1 + np.cos(2. * np.arccos(a) / 3.))
    n = int(n)
    alpha, beta = _engine.empty((n,), dev), _engine.empty((n,), dev)
    _lib.check(L.kgeo_critical_curve(float(a), float(rminus), float(rplus), float(np.sin(th0)), float(np.cos(th0)),
                                     float(np.tan(th0)), n, dptr(alpha), dptr(beta), _engine.stream_ptr()))
    keep = ~(t.isnan(alpha) | t.isnan(beta))          # "fails a lot..." (:292-295): drop NaNs, then mirror
    alpha, beta = alpha[keep], beta[keep]
    alpha = t.cat((alpha, t.flip(alpha, (0,))))
    beta = t.cat((beta, -t.flip(beta, (0,))))
    return (alpha.cpu().numpy(), beta.cpu().numpy())
