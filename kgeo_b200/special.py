"""Device versions of the scipy.special functions the reference calls on the hot path: Carlson
RF/RD/RJ/RC, Legendre K/F/E, Jacobi sn/cn/dn/am.  NumPy in, NumPy out, computed by libkgeo_b200.so kernels."""
from . import _engine


def elliprf(x, y, z):
    return _engine.elementwise("kgeo_special_rf", (x, y, z))


def elliprd(x, y, z):
    return _engine.elementwise("kgeo_special_rd", (x, y, z))


def elliprj(x, y, z, p):
    return _engine.elementwise("kgeo_special_rj", (x, y, z, p))


def elliprc(x, y):
    return _engine.elementwise("kgeo_special_rc", (x, y))


def ellipk(m):
    return _engine.elementwise("kgeo_special_ellipk", (m,))


def ellipkinc(phi, m):
    return _engine.elementwise("kgeo_special_ellipkinc", (phi, m))


def ellipeinc(phi, m):
    return _engine.elementwise("kgeo_special_ellipeinc", (phi, m))


def ellipj(u, m):
    """(sn, cn, dn, am) as scipy.special.ellipj, 0 <= m <= 1"""
    return _engine.elementwise("kgeo_special_ellipj", (u, m), nout=4)
