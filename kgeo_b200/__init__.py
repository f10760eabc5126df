"""kgeo_b200 -- B200-native (sm_100a) analytic Kerr ray tracing: a drop-in for the hot path of achael/kgeo.

Same module and function names as the reference package for the accelerated path:
    kgeo_b200.raytrace_ana, coords_at_tau, Geodesics        (kgeo/kerr_raytracing_ana.py, kerr_raytracing_utils.py)
    kgeo_b200.r_equatorial, nmax_equatorial, nmax_poloidal  (kgeo/equatorial_lensing.py)
    kgeo_b200.make_image, Iobs                              (kgeo/equatorial_images.py)
    kgeo_b200.Velocity, Bfield, Emissivity                  (kgeo/velocities.py, bfields.py, emissivities.py)
Everything is computed by hand-written CUDA kernels in libkgeo_b200.so through a C ABI (include/kgeo_b200.h);
there is no CPU fallback.
"""
from . import _lib
from .kerr_raytracing_utils import *          # noqa: F401,F403
from .kerr_raytracing_ana import *            # noqa: F401,F403
from .equatorial_lensing import *             # noqa: F401,F403
from .equatorial_images import *              # noqa: F401,F403
from .velocities import Velocity              # noqa: F401
from .bfields import Bfield                   # noqa: F401
from .emissivities import Emissivity          # noqa: F401
from . import special, scipy_ellip_binding    # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
