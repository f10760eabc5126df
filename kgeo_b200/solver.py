"""Drop-in for kgeo/solver.py: where null geodesics cross field lines of fixed psi (the jet models' stream functions).

The sign-change scan over the geodesic tracks (`getguesses`), the vectorised secant solve (`findroot`: SciPy's
`optimize.newton` without a derivative, restated on the device including its all-elements-iterate-until-all-converged
termination) and the crossing bookkeeping run in libkgeo_b200.so kernels on the device-resident tracks of a `Geodesics`
object; nothing is evaluated in NumPy here.

psifunc / psi_BZ_* -- kgeo/solver.py:13-37        get_maxtau_forwardjet -- :57-88
getguesses         -- kgeo/solver.py:98-137       findroot              -- :142-207
"""
import ctypes
import warnings

import numpy as np

from . import _engine, _lib
from .kerr_raytracing_utils import MINSPIN

MAXTAUFRAC = (1. - 1.e-10)
PSI_MODELS = dict(para=0, mono=1, power=2)


def _model_id(model):
    return PSI_MODELS.get(model, 3)          # psifunc returns 0 for unknown models (kgeo/solver.py:36-37)


def psi_BZ_para(r, theta, a):
    abscth = np.abs(np.cos(theta))
    rp = 1 + np.sqrt(1 - a ** 2)
    return r * (1 - abscth) + rp * (1 + abscth) * (1 - np.log(1 + abscth)) - 2 * rp * (1 - np.log(2))


def psi_BZ_mono(theta):
    return 1 - np.abs(np.cos(theta))


def psi_power(r, theta, pval):
    return r ** pval * (1 - np.abs(np.cos(theta)))


def psifunc(r, theta, a, model='para', pval=1):
    """the stream functions themselves (closed forms on caller-supplied host arrays, as in the reference)"""
    if model == 'para':
        return psi_BZ_para(r, theta, a)
    elif model == 'mono':
        return psi_BZ_mono(theta)
    elif model == 'power':
        return psi_power(r, theta, pval)
    return 0


def _bundle_of(outgeo, a, r_o, th_o, alpha, beta):
    """the device-resident rays of `outgeo` (made by raytrace_ana), or a fresh per-ray setup for the same pixels"""
    b = getattr(outgeo, '_bundle', None)
    if (b is not None and b.npix == len(alpha) and b.a == float(a) and b.r_o == float(r_o) and b.th_o == float(th_o)
            and b.alpha_host is not None and np.array_equal(b.alpha_host, alpha) and np.array_equal(b.beta_host, beta)):
        return b                             # same observer AND the same pixels (the reference recomputes from alpha, beta)
    return _engine.RayBundle(a, r_o, th_o, np.asarray(alpha, dtype=np.float64), np.asarray(beta, dtype=np.float64),
                             clamp_lowspin=True)


def _guess_call(bundle, outgeo, psitarget, neqmax, model, pval, maxlen, count=None, maxcount=None, guesses=None, taumax=None):
    L = _lib.lib()
    if outgeo is None:                       # no tracks: only the search limit (get_maxtau_forwardjet)
        r = th = tau = None
        ngeo = 0
    else:
        r, th, tau = [_engine.to_device(outgeo.device_array(k), bundle.dev) for k in ('r', 'th', 'tau')]
        ngeo = int(r.shape[0])
    with _engine.torch().cuda.device(bundle.dev):
        _lib.check(L.kgeo_fieldline_guesses(ctypes.byref(bundle.obs), bundle.npix, ngeo, _engine.dptr(r), _engine.dptr(th),
                                            _engine.dptr(tau), bundle.npix, _engine.dptr(bundle.alpha), _engine.dptr(bundle.beta),
                                            _engine.dptr(bundle.w), _engine.iptr(bundle.flags), _model_id(model), float(pval),
                                            float(psitarget), -1 if neqmax is None else int(neqmax), MAXTAUFRAC, int(maxlen),
                                            _engine.iptr(count), _engine.iptr(maxcount), _engine.dptr(guesses),
                                            _engine.dptr(taumax), _engine.stream_ptr()))


def get_maxtau_forwardjet(a, r_o, th_o, alpha, beta, neqmax=1):
    """maximum Mino time before the neqmax-th equatorial crossing (or MAXTAUFRAC * tau_tot)"""
    alpha, beta = np.asarray(alpha, dtype=np.float64), np.asarray(beta, dtype=np.float64)
    bundle = _engine.RayBundle(a, r_o, th_o, alpha, beta, clamp_lowspin=False)
    t = _engine.torch()
    taumax = _engine.empty((bundle.npix,), bundle.dev)
    count = _engine.empty((bundle.npix,), bundle.dev, t.int32)
    maxcount = _engine.empty((1,), bundle.dev, t.int32)
    _guess_call(bundle, None, 1.0, neqmax, 'none', 1, 0, count=count, maxcount=maxcount, taumax=taumax)
    return taumax.cpu().numpy()


def _device_guesses(bundle, outgeo, psitarget, neqmax, model, pval):
    t = _engine.torch()
    count = _engine.empty((bundle.npix,), bundle.dev, t.int32)
    maxcount = _engine.empty((1,), bundle.dev, t.int32)
    _guess_call(bundle, outgeo, psitarget, neqmax, model, pval, 0, count=count, maxcount=maxcount)
    maxlen = int(maxcount.item())
    guesses = _engine.empty((maxlen, bundle.npix), bundle.dev)
    _guess_call(bundle, outgeo, psitarget, neqmax, model, pval, maxlen, guesses=guesses)
    return guesses


def getguesses(outgeo, a, rout, inc, alphas, betas, psitarget, ngeo, neqmax=1, model='para', pval=1):
    """Mino-time guesses of the field-line crossings of every ray: array (max number of crossings, npix), padded with -1"""
    bundle = _bundle_of(outgeo, a, rout, inc, alphas, betas)
    return _device_guesses(bundle, outgeo, psitarget, neqmax, model, pval).cpu().numpy()


def findroot(outgeo, psitarget, alpha, beta, r_o, th_o, a, ngeo, model='para', neqmax=1, tol=1e-8, pval=1):
    """find the crossings with the secant method; returns (tau, r_s, th_s, signpr, signptheta, neqvals, guesses_shape) with the
    per-crossing arrays flattened as in the reference (crossing index slowest)"""
    alpha, beta = np.asarray(alpha, dtype=np.float64), np.asarray(beta, dtype=np.float64)
    bundle = _bundle_of(outgeo, a, r_o, th_o, alpha, beta)
    if bundle.clamped:
        print("WARNING: there were eta<0 points for spin %f<MINSPIN!" % a)
    guesses = _device_guesses(bundle, outgeo, psitarget, neqmax, model, pval)
    guesses_shape = tuple(int(x) for x in guesses.shape)
    print(guesses_shape)
    L = _lib.lib()
    n = guesses_shape[0] * guesses_shape[1]
    outs = [_engine.empty((n,), bundle.dev) for _ in range(6)]
    iters, nfailed = ctypes.c_int32(0), ctypes.c_int32(0)
    print('before solve')
    with _engine.torch().cuda.device(bundle.dev):
        _lib.check(L.kgeo_fieldline_solve(ctypes.byref(bundle.obs), bundle.npix, guesses_shape[0], _engine.dptr(bundle.beta),
                                          _engine.dptr(bundle.w), _engine.iptr(bundle.flags), _model_id(model), float(pval),
                                          float(psitarget), _engine.dptr(guesses), float(tol), 500,
                                          *[_engine.dptr(o) for o in outs], ctypes.byref(iters), ctypes.byref(nfailed),
                                          _engine.stream_ptr()))
    print('after solve')
    # scipy.optimize.newton's array contract (_array_newton): every element failed -> RuntimeError, some -> RuntimeWarning
    # (zero-derivative elements, e.g. the -1 padding whose objective is identically 0, only ever warn there)
    if nfailed.value < 0:
        warnings.warn('%d elements failed to converge after %d iterations (some with zero secant derivative)'
                      % (-nfailed.value, iters.value), RuntimeWarning, stacklevel=2)
    elif nfailed.value >= n:
        raise RuntimeError('convergence failed after %d iterations' % iters.value)
    elif nfailed.value:
        warnings.warn('some failed to converge after %d iterations' % iters.value, RuntimeWarning, stacklevel=2)
    tau, r_s, th_s, signpr, signptheta, neqvals = [o.cpu().numpy() for o in outs]
    return tau, r_s, th_s, signpr, signptheta, neqvals, guesses_shape
