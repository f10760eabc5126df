"""Argument validation shared by the public entry points; same rules and messages as the reference
(kgeo/kerr_raytracing_ana.py:59-71, kgeo/equatorial_images.py:38-45, kgeo/equatorial_lensing.py:94-106)."""
import numpy as np


def check_spin(a):
    if not (isinstance(a, float) and (0 <= np.abs(a) < 1)):
        raise Exception("|a| should be a float in range [0,1)")


def check_observer(a, r_o, th_o, warn=False):
    check_spin(a)
    if warn and a < 0:
        print("WARNING a<0! Not fully verified!")
    if not (isinstance(r_o, float) and (r_o >= 100)):
        raise Exception("r_o should be a float > 100")
    if not (isinstance(th_o, float) and (0 < th_o < np.pi) and th_o != 0.5 * np.pi):
        raise Exception("th_o should be a float in range (0,pi/2) or (pi/2,pi)")
    if warn and th_o > np.pi / 2.:
        print("WARNING th_o>pi/2! Not fully verified!")


def pixel_arrays(alpha, beta, what="alpha, beta are different lengths!"):
    if not isinstance(alpha, np.ndarray):
        alpha = np.array([alpha]).flatten()
    if not isinstance(beta, np.ndarray):
        beta = np.array([beta]).flatten()
    if len(alpha) != len(beta):
        raise Exception(what)
    return np.asarray(alpha, dtype=np.float64), np.asarray(beta, dtype=np.float64)


def require_equatorial(th, what):
    """The element-wise model kernels evaluate in the equatorial plane only; the reference accepts any polar angle."""
    import numpy as np
    th = np.asarray(th, dtype=np.float64)
    if not np.all(th == np.pi / 2):
        raise NotImplementedError("%s: only th = pi/2 (the equatorial plane) is built; off-equatorial evaluation "
                                  "exists on the device for the jet pipeline (off_eq.getstokes) only" % what)
