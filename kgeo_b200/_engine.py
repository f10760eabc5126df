"""Host-side plumbing between NumPy arrays, device buffers (torch CUDA tensors) and the C ABI.

torch is used for device memory, streams and (in sharding.py) torch.distributed only; every computation is a kernel of
libkgeo_b200.so.  Nothing here computes geodesics on the CPU.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import Observer, c_double_p, c_int32_p

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError("kgeo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return t


def device(index=None):
    t = require_cuda()
    return t.device("cuda", t.cuda.current_device() if index is None else index)


def stream_ptr():
    return ctypes.c_void_p(torch().cuda.current_stream().cuda_stream)


def dptr(tensor):
    """device pointer of a float64 CUDA tensor (or NULL for None)"""
    if tensor is None:
        return ctypes.cast(None, c_double_p)
    assert tensor.is_cuda and tensor.dtype == torch().float64 and tensor.is_contiguous()
    return ctypes.cast(tensor.data_ptr(), c_double_p)


def iptr(tensor):
    if tensor is None:
        return ctypes.cast(None, c_int32_p)
    assert tensor.is_cuda and tensor.dtype == torch().int32 and tensor.is_contiguous()
    return ctypes.cast(tensor.data_ptr(), c_int32_p)


def hptr(arr):
    """host pointer of a C-contiguous float64 NumPy array (or NULL for None)"""
    if arr is None:
        return ctypes.cast(None, c_double_p)
    assert arr.dtype == np.float64 and arr.flags.c_contiguous
    return arr.ctypes.data_as(c_double_p)


def to_device(x, dev=None):
    """float64 array-like / tensor -> contiguous float64 CUDA tensor"""
    t = require_cuda()
    if isinstance(x, t.Tensor):
        return x.to(device=dev or device(), dtype=t.float64).contiguous()
    arr = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if not arr.flags.writeable:             # torch.from_numpy warns on read-only arrays (np.load, broadcast views); the copy is host-side only
        arr = arr.copy()
    return t.from_numpy(arr).to(dev or device())


def empty(shape, dev=None, dtype=None):
    t = require_cuda()
    return t.empty(shape, dtype=dtype or t.float64, device=dev or device())


def observer_struct(a, r_o, th_o):
    return Observer(float(a), float(r_o), float(th_o))


def setup_mode(a, th_o, alpha, beta, clamp_lowspin):
    """The reference's whole-array guards np.any(eta>0) / np.any(eta<=0) (kerr_raytracing_ana.py:217,300) and the
    low-spin clamp (:80-82), decided on the host from the input pixel coordinates (no geodesic arithmetic)."""
    alpha = np.asarray(alpha, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    eta = (alpha ** 2 - a ** 2) * np.cos(th_o) ** 2 + beta ** 2
    mode = 0
    clamped = False
    if clamp_lowspin and abs(a) < 1.0e-4 and np.any(eta < 0):
        eta = np.where(eta < 0, 1.0e-12, eta)
        mode |= _lib.MODE_CLAMP_LOWSPIN_ETA
        clamped = True
    if np.any(eta > 0):
        mode |= _lib.MODE_ANY_ETA_POS
    if np.any(eta <= 0):
        mode |= _lib.MODE_ANY_ETA_NONPOS
    if os.environ.get("KGEO_B200_CARLSON_ONLY"):   # A/B switch: the reference's duplication algorithm for every sample
        mode |= _lib.MODE_CARLSON_ONLY
    return mode, clamped


class RayBundle(object):
    """Device-resident rays of one observer: pixel coordinates, the per-ray constant workspace and flags."""

    def __init__(self, a, r_o, th_o, alpha, beta, clamp_lowspin=True, dev=None, mode=None):
        L = _lib.lib()
        self.a, self.r_o, self.th_o = float(a), float(r_o), float(th_o)
        self.obs = observer_struct(a, r_o, th_o)
        self.dev = dev or device()
        if mode is None:
            mode, self.clamped = setup_mode(self.a, self.th_o, alpha, beta, clamp_lowspin)
        else:
            self.clamped = False
        self.mode = mode
        # host copies of the pixel coordinates (when they came from the host): callers that want to reuse this bundle for
        # "the same rays" compare against them (solver._bundle_of)
        self.alpha_host = np.array(alpha, dtype=np.float64) if not hasattr(alpha, "is_cuda") else None
        self.beta_host = np.array(beta, dtype=np.float64) if not hasattr(beta, "is_cuda") else None
        self.alpha = to_device(alpha, self.dev)
        self.beta = to_device(beta, self.dev)
        self.npix = int(self.alpha.numel())
        t = torch()
        self.nw = L.kgeo_ray_workspace_slots()
        with t.cuda.device(self.dev):
            self.w = empty((int(L.kgeo_ray_workspace_doubles(self.npix)),), self.dev)   # grouped layout, see kgeo_b200.h
            self.flags = empty((self.npix,), self.dev, t.int32)
            _lib.check(L.kgeo_ray_setup(ctypes.byref(self.obs), self.npix, dptr(self.alpha), dptr(self.beta), self.mode,
                                        dptr(self.w), iptr(self.flags), stream_ptr()))

    def slot(self, index):
        """one slot of the ray-constant workspace as an [npix] device tensor"""
        t = torch()
        out = empty((self.npix,), self.dev)
        with t.cuda.device(self.dev):
            _lib.check(_lib.lib().kgeo_ray_slot(self.npix, dptr(self.w), int(index), dptr(out), stream_ptr()))
        return out

    @property
    def tau_tot(self):
        return self.slot(2)       # slot S_TAUTOT

    def tracks(self, ntau=None, maxtaufrac=1.0, tau=None, full=True, want=("t", "r", "th", "ph", "sig", "tau")):
        """Launch the track kernel.  tau: None (linspace samples) or a [ntau, npix] device tensor.  full: False = r, theta
        only; True = all coordinates (hot + fix-up launches); 2 = all coordinates through the one-pass complete kernel.
        Returns a dict name -> [ntau, npix] float64 CUDA tensor for the names in `want`."""
        L = _lib.lib()
        t = torch()
        if tau is not None:
            tau = to_device(tau, self.dev).reshape(-1, self.npix)
            ntau = int(tau.shape[0])
        out = {}
        with t.cuda.device(self.dev):
            for name in ("t", "r", "th", "ph", "sig", "tau"):
                out[name] = empty((ntau, self.npix), self.dev) if name in want else None
            _lib.check(L.kgeo_tracks(ctypes.byref(self.obs), self.npix, dptr(self.w), iptr(self.flags), ntau,
                                     float(maxtaufrac), dptr(tau), self.npix, (2 if full == 2 else int(bool(full))),
                                     dptr(out["t"]), dptr(out["r"]), dptr(out["th"]), dptr(out["ph"]), dptr(out["sig"]),
                                     dptr(out["tau"]), self.npix, stream_ptr()))
        return {k: v for k, v in out.items() if v is not None}

    def tracks_host(self, ntau, maxtaufrac=1.0, want=("t", "r", "th", "ph")):
        """The host pipeline of the C ABI (kgeo_raytrace_host) for the rays of this bundle: setup + track kernels in sample
        slabs overlapped with D2H copies into page-locked NumPy arrays (torch's caching pinned allocator).  Returns a dict
        name -> (ntau, npix) float64 NumPy array for the names in `want`.  Bit-identical to `tracks` (same kernels)."""
        L = _lib.lib()
        t = torch()
        host = {}
        for name in ("t", "r", "th", "ph", "sig", "tau"):
            if name in want:
                try:
                    host[name] = t.empty((ntau, self.npix), dtype=t.float64, pin_memory=True)
                except RuntimeError:                      # page-locked memory exhausted: pageable destination
                    host[name] = t.empty((ntau, self.npix), dtype=t.float64)
        hp = lambda n: ctypes.cast(host[n].data_ptr(), c_double_p) if n in host else ctypes.cast(None, c_double_p)
        _lib.check(L.kgeo_raytrace_host(ctypes.byref(self.obs), self.npix, hptr(self.alpha_host), hptr(self.beta_host), self.mode,
                                        int(ntau), float(maxtaufrac), hp("t"), hp("r"), hp("th"), hp("ph"), hp("sig"), hp("tau"),
                                        self.dev.index if self.dev.index is not None else t.cuda.current_device()))
        return {k: v.numpy() for k, v in host.items()}        # the arrays keep the pinned blocks alive (arr.base)

    def crossing_counts(self, tau, want_eq=True, want_pol=True):
        L = _lib.lib()
        t = torch()
        tau = to_device(tau, self.dev).reshape(-1, self.npix)
        ntau = int(tau.shape[0])
        with t.cuda.device(self.dev):
            neq = empty((ntau, self.npix), self.dev) if want_eq else None
            npol = empty((ntau, self.npix), self.dev) if want_pol else None
            _lib.check(L.kgeo_crossing_counts(ctypes.byref(self.obs), self.npix, dptr(self.beta), dptr(self.w),
                                              iptr(self.flags), ntau, dptr(tau), dptr(neq), dptr(npol), stream_ptr()))
        return neq, npol


_SUMMARY_KEYS = ("tau_tot", "uplus", "uminus", "uratio", "a2uminus", "outside", "thplus", "thminus")


def _summary(a, r_o, th_o, x, y, mode, want):
    L = _lib.lib()
    t = require_cuda()
    dev = device()
    obs = observer_struct(a, r_o, th_o)
    dx, dy = to_device(x, dev), to_device(y, dev)
    n = int(dx.numel())
    out = dict(roots=empty((8, n), dev) if "roots" in want else None,
               rclass=empty((n,), dev, t.int32) if "rclass" in want else None)
    for k in _SUMMARY_KEYS:
        out[k] = empty((n,), dev) if k in want else None
    _lib.check(L.kgeo_ray_summary(ctypes.byref(obs), n, dptr(dx), dptr(dy), mode, dptr(out["roots"]), iptr(out["rclass"]),
                                  *[dptr(out[k]) for k in _SUMMARY_KEYS], stream_ptr()))
    return {k: v for k, v in out.items() if v is not None}


def ray_summary(a, r_o, th_o, alpha, beta, clamp_lowspin=False, want=("roots", "rclass", "tau_tot")):
    """kgeo_ray_summary on pixel coordinates -> dict of device tensors."""
    mode, _ = setup_mode(a, th_o, alpha, beta, clamp_lowspin)
    return _summary(a, r_o, th_o, alpha, beta, mode, want)


def ray_summary_lam_eta(a, r_o, th_o, lam, eta, want=("roots", "rclass")):
    """kgeo_ray_summary on conserved quantities (lam, eta) -> dict of device tensors."""
    return _summary(a, r_o, th_o, lam, eta, _lib.MODE_INPUT_LAM_ETA, want)


def mino_total_from_roots(a, r_o, eta, roots):
    """kgeo_mino_total: roots is an [8, n] array (Re r1..r4, Im r1..r4)."""
    L = _lib.lib()
    require_cuda()
    dev = device()
    obs = observer_struct(a, r_o, 0.25 * np.pi)
    droots = to_device(roots, dev)
    n = int(droots.shape[1])
    out = empty((n,), dev)
    _lib.check(L.kgeo_mino_total(ctypes.byref(obs), n, dptr(droots), dptr(out), stream_ptr()))
    return out


def elementwise(fn_name, inputs, nout=1):
    """Run an element-wise special-function kernel on broadcast NumPy inputs; returns NumPy array(s)."""
    L = _lib.lib()
    require_cuda()
    arrs = np.broadcast_arrays(*[np.asarray(x, dtype=np.float64) for x in inputs])
    shape = arrs[0].shape
    dev = device()
    dins = [to_device(np.ascontiguousarray(a).ravel(), dev) for a in arrs]
    n = int(dins[0].numel())
    outs = [empty((n,), dev) for _ in range(nout)]
    _lib.check(getattr(L, fn_name)(n, *[dptr(d) for d in dins], *[dptr(o) for o in outs], stream_ptr()))
    res = [o.cpu().numpy().reshape(shape) for o in outs]
    return res[0] if nout == 1 else tuple(res)
