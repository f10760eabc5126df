"""Drop-in for kgeo/kerr_raytracing_ana.py: raytrace_ana / coords_at_tau with the reference's signatures, argument
checks, messages and outputs, computed by the fused sm_100a track kernel (libkgeo_b200.so).

raytrace_ana  -- kgeo/kerr_raytracing_ana.py:46-122
coords_at_tau -- kgeo/kerr_raytracing_ana.py:126-193
"""
import time

import numpy as np

from . import _checks, _engine
from .kerr_raytracing_utils import Geodesics, MINSPIN

# defaults of the reference (kgeo/kerr_raytracing_ana.py:21-29)
SPIN = 0.94
INC = 20 * np.pi / 180.
ROUT = 1.e5
NGEO = 250
NPIX = 1000
MAXTAUFRAC = (1. - 1.e-10)

alpha_default = np.linspace(-6, 6, NPIX)
beta_default = 0 * alpha_default

# the reference's module-level backend switch (kgeo/kerr_raytracing_ana.py:31-43): here the only backend is the device one
SCIPY = False
GSL = False
B200 = True


def _track_names(npix, ngeo, dev):
    """Which of the six (ngeo, npix) arrays to materialise eagerly: all of them when they fit comfortably in HBM,
    otherwise only t, r, theta, phi (sigma and tau are then produced on first access)."""
    t = _engine.torch()
    free, _total = t.cuda.mem_get_info(dev)
    need6 = 6 * 8 * npix * ngeo
    if need6 < 0.6 * free:
        return ("t", "r", "th", "ph", "sig", "tau")
    return ("t", "r", "th", "ph")


def raytrace_ana(a=SPIN,
                 observer_coords=[0, ROUT, INC, 0],
                 image_coords=[alpha_default, beta_default],
                 ngeo=NGEO, maxtaufrac=MAXTAUFRAC,
                 do_phi_and_t=True,
                 savedata=False, plotdata=True):
    """analytically raytrace over the full range of mino times"""
    tstart = time.time()

    [_, r_o, th_o, _] = observer_coords  # assumes ph_o = 0
    [alpha, beta] = image_coords

    _checks.check_observer(a, r_o, th_o, warn=True)
    alpha, beta = _checks.pixel_arrays(alpha, beta)

    print('calculating preliminaries...')
    bundle = _engine.RayBundle(a, r_o, th_o, alpha, beta, clamp_lowspin=True)
    if bundle.clamped:
        print("WARNING: there were eta<0 points for spin %f<MINSPIN!" % a)

    # The (ngeo, npix) arrays are produced on first use, in the form the caller asks for:
    #  * NumPy fields (geo_coords, r_s, mino_times, ...) -> the host pipeline of the C ABI (kgeo_raytrace_host): the track kernels
    #    run in sample slabs while the previous slab drains into page-locked NumPy arrays, so a script that only reads the host
    #    arrays gets them at the PCIe rate and the 16.8 GB of a 1024^2 x 500 trace never sit in HBM at once;
    #  * device_array(name) (the solver, sharding, anything that stays on the GPU) -> one launch into device-resident tensors.
    # As in the reference (:97) phi and t are always computed, whatever do_phi_and_t says.
    def device_tracks(want):
        names = set(want)
        if names & {"t", "r", "th", "ph"}:            # one launch gives everything: take sigma and tau along when they fit
            names |= set(_track_names(bundle.npix, ngeo, bundle.dev))
        return bundle.tracks(ntau=ngeo, maxtaufrac=maxtaufrac, full=True, want=tuple(names))

    def host_tracks(want):
        return bundle.tracks_host(ntau=ngeo, maxtaufrac=maxtaufrac, want=tuple(want))

    geos = Geodesics(a, observer_coords, [alpha, beta], None, None, [None, None, None, None], _bundle=bundle,
                     _device_tracks=device_tracks, _host_tracks=host_tracks if bundle.alpha_host is not None else None)

    if savedata and do_phi_and_t:      # the reference wraps this in try / except and carries on (:103-108)
        print('saving data...')
        try:
            geos.savegeos()
        except ImportError as exc:
            print("not saved: %s" % exc)
    if plotdata and do_phi_and_t:
        print("plotdata=True ignored: plotgeos is not built (matplotlib host-side code, SURVEY 2 out of scope)")

    _engine.torch().cuda.synchronize(bundle.dev)
    tstop = time.time()
    print('done!  ', tstop - tstart, ' seconds!')
    return geos


def coords_at_tau(a, observer_coords, image_coords, tau, do_phi_and_t=True):
    """compute coordinates x^mu at a particular Mino time: returns (sig_s, [t_s, r_s, th_s, ph_s]) as NumPy arrays
    of shape (ntau, npix)"""
    [_, r_o, th_o, _] = observer_coords  # assumes ph_o = 0
    [alpha, beta] = image_coords

    _checks.check_observer(a, r_o, th_o, warn=True)
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    is_tensor = hasattr(tau, "is_cuda")
    if not is_tensor:
        tau = np.asarray(tau, dtype=np.float64)
    if len(tau.shape) == 1:
        tau = tau.reshape(1, tau.shape[0])
    if not (tau.shape[1] == len(alpha)):
        raise Exception("tau has incompatible shape in coords_at_tau!")

    bundle = _engine.RayBundle(a, r_o, th_o, alpha, beta, clamp_lowspin=True)
    if bundle.clamped:
        print("WARNING: there were eta<0 points for spin %f<MINSPIN!" % a)
    out = bundle.tracks(tau=tau, full=bool(do_phi_and_t), want=("t", "r", "th", "ph", "sig"))
    host = {k: v.cpu().numpy() for k, v in out.items()}
    return (host["sig"], [host["t"], host["r"], host["th"], host["ph"]])


def _integrals(a, r_o, th_o, lam, eta, s_o, tausteps, do_phi_and_t, side):
    """Rays set up on the device from the conserved quantities (kgeo_ray_setup_conserved), then the polar or the radial
    antiderivatives at `tausteps` (kgeo_track_integrals).  Returns NumPy arrays of tausteps' shape."""
    import ctypes
    from . import _lib
    L = _lib.lib()
    t = _engine.require_cuda()
    dev = _engine.device()
    shape = tausteps.shape
    npix = len(eta)
    tau2 = np.ascontiguousarray(np.asarray(tausteps, dtype=np.float64).reshape(-1, npix))
    ntau = tau2.shape[0]
    mode = (_lib.MODE_ANY_ETA_POS if np.any(eta > 0.) else 0) | (_lib.MODE_ANY_ETA_NONPOS if np.any(eta <= 0.) else 0)
    obs = _engine.observer_struct(a, r_o, th_o)
    d = lambda x: _engine.to_device(np.ascontiguousarray(x, dtype=np.float64), dev)
    d_lam, d_eta, d_tau = d(lam), d(eta), d(tau2)
    d_s = d(s_o) if s_o is not None else None
    names = ("th", "Gph", "Gt") if side == "polar" else ("r", "Iph", "It", "Isig")
    with t.cuda.device(dev):
        w = _engine.empty((int(L.kgeo_ray_workspace_doubles(npix)),), dev)
        flags = _engine.empty((npix,), dev, t.int32)
        _lib.check(L.kgeo_ray_setup_conserved(ctypes.byref(obs), npix, _engine.dptr(d_lam), _engine.dptr(d_eta),
                                              _engine.dptr(d_s), mode, int(bool(do_phi_and_t)), _engine.dptr(w),
                                              _engine.iptr(flags), _engine.stream_ptr()))
        out = {k: _engine.empty((ntau, npix), dev) for k in names}
        p = lambda k: _engine.dptr(out.get(k))
        _lib.check(L.kgeo_track_integrals(ctypes.byref(obs), npix, _engine.dptr(w), _engine.iptr(flags), ntau,
                                          _engine.dptr(d_tau), npix, int(bool(do_phi_and_t)), p("th"), p("Gph"), p("Gt"),
                                          p("r"), p("Iph"), p("It"), p("Isig"), npix, _engine.stream_ptr()))
    return tuple(out[k].cpu().numpy().reshape(shape) for k in names)


def _as1d(x):
    return x if isinstance(x, np.ndarray) else np.array([x]).flatten()


def th_integrate(a, th_o, s_o, lam, eta, u_plus, u_minus, tausteps, do_phi_and_t=True):
    """(th_s, G_ph, G_t) at the Mino times `tausteps` (shape (..., npix)) for rays given by their conserved quantities and the
    sign s_o of the observer's polar momentum (kgeo/kerr_raytracing_ana.py:195-365; same checks and messages).  u_plus / u_minus
    are the turning points uplus_uminus returns for (lam, eta); the device setup recomputes them from (lam, eta), so arrays
    that are NOT those of uplus_uminus are not honoured."""
    if not (isinstance(a, float) and (0 <= np.abs(a) < 1)):
        raise Exception("|a| should be a float in range [0,1)")
    if not (isinstance(th_o, float) and (0 < th_o < np.pi) and th_o != 0.5 * np.pi):
        raise Exception("th_o should be a float in range (0,pi/2) or (pi/2,pi)")
    s_o, eta, lam, u_plus, u_minus = [_as1d(x) for x in (s_o, eta, lam, u_plus, u_minus)]
    if not (len(s_o) == len(eta) == len(lam) == len(u_plus) == len(u_minus)):
        raise Exception("inputs to th_integrate not the same length!")
    if not (tausteps.shape[-1] == len(s_o)):
        raise Exception("tausteps has incompatible shape in th_integrate!")
    if np.any(eta <= 0.) and np.abs(a) < 1.e-4 and np.any(eta < 0.):
        raise Exception("below MINSPIN but there are eta<0 points in th_integrate!")
    return _integrals(a, 1000., th_o, lam, eta, s_o, tausteps, do_phi_and_t, "polar")   # the polar side does not depend on r_o


def r_integrate(a, r_o, lam, eta, r1, r2, r3, r4, tausteps, do_phi_and_t=True):
    """(r_s, I_ph, I_t, I_sig) at the Mino times `tausteps` (kgeo/kerr_raytracing_ana.py:368-633; same checks and messages).
    r1..r4 are the roots radial_roots returns for (lam, eta); the device setup recomputes them.  r_o = inf needs
    do_phi_and_t=False (the reference's phi, t integrals diverge there; r_equatorial is the caller that uses it)."""
    if not (isinstance(a, float) and (0 <= np.abs(a) < 1)):
        raise Exception("|a| should be a float in range [0,1)")
    if not (isinstance(r_o, float) and (r_o >= 100)):
        raise Exception("r_o should be a float > 100")
    lam, eta, r1, r2, r3, r4 = [_as1d(x) for x in (lam, eta, r1, r2, r3, r4)]
    if not (len(lam) == len(eta) == len(r1) == len(r2) == len(r3) == len(r4)):
        raise Exception("inputs to r_integrate not the same length!")
    if not (tausteps.shape[-1] == len(eta)):
        raise Exception("tausteps has incompatible shape in r_integrate!")
    # th_o only enters the polar side; any admissible value serves the radial setup
    return _integrals(a, r_o, 0.3, lam, eta, None, tausteps, do_phi_and_t, "radial")


def _legendre(phi, j):
    phi, j = np.broadcast_arrays(np.asarray(phi, dtype=np.float64), np.asarray(j, dtype=np.float64))
    return (_engine.elementwise("kgeo_special_ellipkinc", (phi, j)), _engine.elementwise("kgeo_special_ellipeinc", (phi, j)))


def S1_S2(al, phi, j, ret_s2=True):
    """GL19a (B92), (B93): the case-4 radial antiderivatives S1, S2 of amplitude phi, parameter j and offset al
    (kgeo/kerr_raytracing_ana.py:638-663), the elliptic integrals from the device library."""
    from .scipy_ellip_binding import ellip_pi_arr
    al, phi, j = [np.asarray(x, dtype=np.float64) for x in (al, phi, j)]
    n1 = 1. + al * al                                  # 1 + al^2: the characteristic
    nj = n1 - j
    delta = np.sqrt(1. - j * np.sin(phi) ** 2)
    p2 = np.sqrt(n1 / nj)
    f2 = 0.5 * p2 * np.log(np.abs(((1. - p2) * (1. + p2 * delta)) / ((1. + p2) * (1. - p2 * delta))))
    F, E = _legendre(phi, j)
    S1 = (F + al * al * ellip_pi_arr(n1, phi, j) - al * f2) / n1
    if not ret_s2:
        return (S1, np.nan)
    tp = np.tan(phi)
    first = -((1. - j) * F + al * al * (E + delta * (al - tp) / (1. + al * tp) - al)) / (n1 * nj)
    return (S1, first + S1 * (1. / n1 + (1. - j) / nj))


def R1_R2(al, phi, j, ret_r2=True):
    """GL19a (B62), (B65): the case-3 radial antiderivatives R1, R2 (kgeo/kerr_raytracing_ana.py:666-683)."""
    from .scipy_ellip_binding import ellip_pi_arr
    al, phi, j = [np.asarray(x, dtype=np.float64) for x in (al, phi, j)]
    al2 = al * al
    sphi = np.sin(phi)
    delta = np.sqrt(1. - j * sphi ** 2)
    den = j + (1. - j) * al2
    p1 = np.sqrt((al2 - 1.) / den)
    f1 = 0.5 * p1 * np.log(np.abs((p1 * delta + sphi) / (p1 * delta - sphi)))
    nn = al2 / (al2 - 1.)
    R1 = (ellip_pi_arr(nn, phi, j) - al * f1) / (1. - al2)
    if not ret_r2:
        return (R1, np.nan)
    F, E = _legendre(phi, j)
    R2 = (F - (al2 / den) * (E - al * sphi * delta / (1. + al * np.cos(phi)))) / (al2 - 1.) + (2. * j - nn) * R1 / den
    return (R1, R2)
