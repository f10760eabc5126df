"""Drop-in for the hot-path part of kgeo/equatorial_lensing.py: r_equatorial (:85-196), nmax_equatorial (:22-51),
nmax_poloidal (:53-82), computed by libkgeo_b200.so.

Out of scope: rho_of_req*, critical_curve, plot_curves, generate_library (serial Brent root-finding callers).
"""
import ctypes

import numpy as np

from . import _checks, _engine, _lib
from ._engine import dptr

RMAX = np.inf  # kgeo/equatorial_lensing.py:19


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
