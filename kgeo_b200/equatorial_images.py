"""Drop-in for kgeo/equatorial_images.py: make_image (:23-122) and Iobs (:124-264) with the reference's signatures and
return tuples, computed by ONE library call per image (all orders: a geometry and an emission launch; no geodesic array goes
through HBM).
"""
import ctypes
import time

import numpy as np

from . import _checks, _engine, _lib
from ._engine import dptr
from .bfields import Bfield
from .emissivities import Emissivity
from .models import pack_models
from .velocities import Velocity

bfield_default = Bfield('rad')
vel_default = Velocity('zamo')
emis_default = Emissivity('bpl')
SPECIND = 1       # default (negative) spectral index
DISKANGLE = 0.1   # default h/r for equatorial disk
SINTHB = np.pi / 4.  # assumption for sin(thetab) if polarization==False
OBSFREQ = 230.e9  # default observation frequency

_PLANES = ("I", "Q", "U", "r", "Ir", "g", "sinthb")


def pixel_grid(alpha_min, alpha_max, beta_min, beta_max, psize):
    """The reference's pixel grid, quirks included (kgeo/equatorial_images.py:48-55)."""
    n_alpha = int(np.floor(alpha_max - alpha_min) / psize)
    alphas = np.linspace(alpha_min, alpha_min + n_alpha * psize, n_alpha)
    n_beta = int(np.floor(beta_max - beta_min) / psize)
    betas = np.linspace(beta_min, beta_min + n_beta * psize, n_beta)
    alpha_arr, beta_arr = np.meshgrid(alphas, betas)
    return alpha_arr.flatten(), beta_arr.flatten()


def image_device(a, r_o, th_o, mbar_min, mbar_max, alpha, beta, emissivity, velocity, bfield, polarization, pathlength,
                 specind, diskangle, nu_obs, dev=None):
    """Launch the image kernels (all orders in one call); alpha/beta may be NumPy arrays or CUDA tensors.  Returns a dict of CUDA tensors:
    planes [norder, npix] for I,Q,U,r,Ir,g,sinthb(,lp,Ie) and [npix] for nmax, npol, imax."""
    L = _lib.lib()
    _engine.require_cuda()
    dev = dev or _engine.device()
    obs = _engine.observer_struct(a, r_o, th_o)
    M = pack_models(emissivity=emissivity, velocity=velocity, bfield=bfield)
    opts = _lib.ImageOpts(int(mbar_min), int(mbar_max), int(bool(polarization)), int(bool(pathlength)), float(specind),
                          float(diskangle), float(nu_obs))
    al, be = _engine.to_device(alpha, dev), _engine.to_device(beta, dev)
    n = int(al.numel())
    norder = mbar_max - mbar_min + 1
    t = _engine.torch()
    with t.cuda.device(dev):
        out = {k: _engine.empty((norder, n), dev) for k in _PLANES}
        for k in ("nmax", "npol", "imax"):
            out[k] = _engine.empty((n,), dev)
        out["lp"] = _engine.empty((norder, n), dev) if pathlength else None
        out["Ie"] = _engine.empty((norder, n), dev) if pathlength else None
        _lib.check(L.kgeo_equatorial_image(ctypes.byref(obs), ctypes.byref(M), ctypes.byref(opts), n, dptr(al), dptr(be),
                                           *[dptr(out[k]) for k in _PLANES], dptr(out["nmax"]), dptr(out["npol"]),
                                           dptr(out["imax"]), dptr(out["lp"]), dptr(out["Ie"]), _engine.stream_ptr()))
    return out


_DST_ORDER = _PLANES + ("nmax", "npol", "imax", "lp", "Ie")


def image_into(a, r_o, th_o, mbar_min, mbar_max, alpha, beta, emissivity, velocity, bfield, polarization, pathlength,
               specind, diskangle, nu_obs, dst, cols, row_step, row_offset, plane_stride, dev=None):
    """The image kernels for a shard of whole pixel rows (alpha/beta: NumPy arrays or CUDA tensors, `cols` pixels per row),
    every result stored at its place in DESTINATION planes (kgeo_equatorial_image_sharded): `dst` maps an output name
    ('I','Q','U','r','Ir','g','sinthb','nmax','npol','imax','lp','Ie') to the device address (int) of its first plane --
    memory of this GPU or of another GPU of the box mapped by sharding.PeerBuffer; local row lr lands in row
    lr * row_step + row_offset, order planes are plane_stride elements apart.  Runs on the current stream; returns nothing."""
    L = _lib.lib()
    _engine.require_cuda()
    dev = dev or _engine.device()
    obs = _engine.observer_struct(a, r_o, th_o)
    M = pack_models(emissivity=emissivity, velocity=velocity, bfield=bfield)
    opts = _lib.ImageOpts(int(mbar_min), int(mbar_max), int(bool(polarization)), int(bool(pathlength)), float(specind),
                          float(diskangle), float(nu_obs))
    al, be = _engine.to_device(alpha, dev), _engine.to_device(beta, dev)
    smap = _lib.ShardMap(int(cols), int(row_step), int(row_offset), int(plane_stride), 0)
    ptr = lambda k: ctypes.cast(ctypes.c_void_p(dst.get(k) or None), _lib.c_double_p)
    t = _engine.torch()
    with t.cuda.device(dev):
        _lib.check(L.kgeo_equatorial_image_sharded(ctypes.byref(obs), ctypes.byref(M), ctypes.byref(opts), int(al.numel()),
                                                   dptr(al), dptr(be), ctypes.byref(smap), *[ptr(k) for k in _DST_ORDER],
                                                   _engine.stream_ptr()))
        al.record_stream(t.cuda.current_stream(dev))
        be.record_stream(t.cuda.current_stream(dev))


def make_image(a, r_o, th_o, mbar_max, alpha_min, alpha_max, beta_min, beta_max, psize,
               nmax_only=False,
               emissivity=emis_default,
               bfield=bfield_default,
               velocity=vel_default,
               polarization=False,
               pathlength=False,
               specind=SPECIND,
               diskangle=DISKANGLE,
               nu_obs=OBSFREQ):
    """computes an image in range (alpha_min, alpha_max) x (beta_min, beta_max)
      for all orders of m up to mbar_max and pixel size psize.
      Returns (I, Q, U, r, t, g, sinthb, n, np[, lp, Ie]); the 2-D arrays are indexed [pixel, mbar]."""
    _checks.check_observer(a, r_o, th_o)
    if not (isinstance(mbar_max, int) and (mbar_max >= 0)):
        raise Exception("mbar_max should be an integer >=0!")

    alpha_arr, beta_arr = pixel_grid(alpha_min, alpha_max, beta_min, beta_max, psize)

    if nmax_only:
        # the reference falls off the end here (outdat is never bound, equatorial_images.py:72-79,122)
        raise UnboundLocalError("cannot access local variable 'outdat' where it is not associated with a value")

    tstart = time.time()
    out = image_device(a, r_o, th_o, 0, mbar_max, alpha_arr, beta_arr, emissivity, velocity, bfield, polarization,
                       pathlength, specind, diskangle, nu_obs)
    host = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
    for mbar in range(mbar_max + 1):
        print('image %i...%0.2f s' % (mbar, (time.time() - tstart) / (mbar_max + 1)))

    # planes are order-major on the device; expose the reference's (npix, mbar_max+1) indexing
    res = tuple(host[k].T for k in ("I", "Q", "U", "r", "Ir", "g", "sinthb")) + (host["nmax"], host["npol"])
    if pathlength:
        res = res + (host["lp"].T, host["Ie"].T)
    return res


def Iobs(a, r_o, th_o, mbar, alpha, beta,
         emissivity=emis_default, velocity=vel_default, bfield=bfield_default,
         polarization=False, pathlength=False, specind=SPECIND, diskangle=DISKANGLE, nu_obs=OBSFREQ,
         th_s=np.pi / 2):
    """Return (Iobs, Qobs, Uobs, g, r_s, sin_thb, Ir, Imax, Nmax[, lp, Ie]) for the ring of order mbar"""
    _checks.check_observer(a, r_o, th_o)
    if not (isinstance(mbar, int) and (mbar >= 0)):
        raise Exception("mbar should be an integer >=0!")
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    if th_s != np.pi / 2:
        raise NotImplementedError("only the equatorial plane th_s = pi/2 is accelerated")
    out = image_device(a, r_o, th_o, mbar, mbar, alpha, beta, emissivity, velocity, bfield, polarization, pathlength,
                       specind, diskangle, nu_obs)
    h = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
    if not np.any(h["I"][0] != 0) and not np.any(h["g"][0] != 0):
        print("masked all pixels in Iobs! m=%i" % mbar)
    res = (h["I"][0], h["Q"][0], h["U"][0], h["g"][0], h["r"][0], h["sinthb"][0], h["Ir"][0], h["imax"], h["nmax"])
    if pathlength:
        res = res + (h["lp"][0], h["Ie"][0])
    return res
