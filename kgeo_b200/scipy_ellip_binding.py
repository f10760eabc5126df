"""Drop-in for kgeo/scipy_ellip_binding.py: the incomplete elliptic integral of the third kind Pi(n; phi | m) in the
A&S / Mathematica convention, with the reference's special-value rules (:125-162), evaluated on the device with fp64
Carlson RF/RJ duplication (libkgeo_b200.so) instead of scipy.special.elliprf/elliprj/elliprc."""
import numpy as np

from . import _engine


def ellip_pi_arr(n, phi, m):
    """Vectorised Pi(n; phi | m); n and m are per-ray (length npix), phi is (npix,) or (ngeo, npix)
    (kgeo/scipy_ellip_binding.py:87-237)."""
    if not (isinstance(phi, np.ndarray)):
        phi = np.array([phi]).flatten()
    if not (isinstance(m, np.ndarray)):
        m = np.array([m]).flatten()
    if not (isinstance(n, np.ndarray)):
        n = np.array([n]).flatten()
    if len(phi.shape) > 1:
        n = np.broadcast_to(n, phi.shape)
        m = np.broadcast_to(m, phi.shape)
    return _engine.elementwise("kgeo_special_ellippi", (n, phi, m))


def ellip_pi(n, phi, m):
    """Scalar Pi(n; phi | m) (kgeo/scipy_ellip_binding.py:9-85)."""
    return float(_engine.elementwise("kgeo_special_ellippi", ([n], [phi], [m]))[0])
