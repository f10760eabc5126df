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
