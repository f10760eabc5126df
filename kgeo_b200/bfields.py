"""Magnetic field model descriptor with the constructor / bfield_lab / omega_field / efield_lab / bsq interface of
kgeo/bfields.py:33-131,175-235,365-396.  Every lab-frame model runs on the device (kg_image.cuh), including the perturbative
BZ monopole (linear spline of f(r)), the BZ paraboloid (Lambert W) and the power-law stream functions.  'const_comoving' is
not offered: in the reference itself every image path raises for it (bfield_lab refuses comoving fields and
calc_polquantities uses `norm_x` before assignment, kgeo/equatorial_images.py:422-436)."""
import ctypes

import numpy as np

from . import _checks, _engine, _lib

RMAXINTERP = 100       # kgeo/bfields.py:16-20: bz_fr_data.dat holds f(r) on linspace(1, RMAXINTERP, NINTERP)
NINTERP = 100000
_FR_TABLE = None


def _li2_unit(x):
    """Real dilogarithm Li2(x) for 0 <= x <= 1 (power series below 1/2, Euler's reflection above)."""
    x = np.asarray(x, dtype=np.float64)
    small = np.where(x <= 0.5, x, 1.0 - x)
    k = np.arange(1, 61, dtype=np.float64)
    ser = np.zeros_like(small)
    term = np.ones_like(small)
    for kk in k:
        term = term * small
        ser = ser + term / (kk * kk)
    with np.errstate(divide='ignore', invalid='ignore'):
        refl = np.pi ** 2 / 6. - np.where((x > 0) & (x < 1), np.log(x) * np.log1p(-x), 0.0) - ser
    return np.where(x <= 0.5, ser, refl)


def bz_fr(r):
    """f(r) of the perturbative BZ monopole: the real part of the reference's complex formula (kgeo/bfields.py:7-13),
    with Re Li2(x) = pi^2/3 - ln^2(x)/2 - Li2(1/x) and Re ln(1 - x) = ln|1 - x| for x = 2/r > 1."""
    r = np.asarray(r, dtype=np.float64)
    x = 2. / r
    with np.errstate(divide='ignore', invalid='ignore'):
        li = np.where(x <= 1, _li2_unit(np.minimum(x, 1.0)),
                      np.pi ** 2 / 3. - 0.5 * np.log(x) ** 2 - _li2_unit(np.minimum(1. / x, 1.0)))
        out = (li - np.log(np.abs(1 - x)) * np.log(r / 2.)) * r * r * (2 * r - 3) / 8.
        out += (1 + 3 * r - 6 * r * r) * np.log(r / 2.) / 12.
        out += 11. / 72. + 1. / (3. * r) + r / 2. - r * r / 2.
    return out


def bz_fr_table():
    """(rsinterp, fsinterp) of the reference's bz_fr_data.dat, rebuilt from its generating formula (the reference tree
    does not ship the file; kgeo/bfields.py:15-24 shows how it was made)."""
    global _FR_TABLE
    if _FR_TABLE is None:
        rs = np.linspace(1, RMAXINTERP, NINTERP)
        _FR_TABLE = (rs, bz_fr(rs))
    return _FR_TABLE


class Bfield(object):
    """b-field as a function of r in the equatorial plane"""

    def __init__(self, fieldtype="rad", **kwargs):
        self.fieldtype = fieldtype
        if self.fieldtype in ['rad', 'vert', 'tor', 'simple', 'simple_rm1', 'monopoleA', 'bz_monopole', 'bz_guess',
                              'bz_para', 'power', 'fromfile', 'gen_power']:
            self.fieldframe = 'lab'
        elif self.fieldtype in ['const_comoving']:
            raise NotImplementedError("Bfield('const_comoving'): every image path of the reference raises for comoving "
                                      "fields (kgeo/bfields.py:100-101, kgeo/equatorial_images.py:422-436)")
        else:
            raise Exception("fieldtype %s not recognized in Bfield!" % self.fieldtype)
        self.kwargs = kwargs

        self._table = None
        if self.fieldtype in ['bz_monopole', 'bz_guess', 'bz_para']:
            self.secondorder_only = self.kwargs.get('secondorder_only', False)
            self.C = self.kwargs.get('C', 1)
            if self.fieldtype == 'bz_monopole':
                # extension: fr_table=(r, f) replaces the rebuilt bz_fr_data.dat (e.g. the original file, if at hand)
                self.fr_table = self.kwargs.get('fr_table', None)
        elif self.fieldtype == 'power':
            self.secondorder_only = self.kwargs.get('secondorder_only', False)
            self.C = self.kwargs.get('C', 1)
            self.pval = self.kwargs.get('p', 1)
            self.usemono = self.kwargs.get('usemono', False)
        elif self.fieldtype in ['monopoleA']:
            self.omegafac = self.kwargs.get('omegafac', 0.5)
            self.C = self.kwargs.get('C', 1)
        elif self.fieldtype == 'rad':
            self.Cr, self.Cvert, self.Cph = 1, 0, 0
        elif self.fieldtype == 'vert':
            self.Cr, self.Cvert, self.Cph = 0, 1, 0
        elif self.fieldtype == 'tor':
            self.Cr, self.Cvert, self.Cph = 0, 0, 1
        elif self.fieldtype == 'fromfile':
            from .models import load_cache_from_file
            self.filename = self.kwargs.get('file', None)
            radii, Br, Bth, Bph = load_cache_from_file(self.filename, ('Br', 'Btheta', 'Bphi'))
            self.cached_data = dict(radii=radii, Br=Br, Bth=Bth, Bph=Bph)
            self._table = None
        elif self.fieldtype == 'gen_power':
            self.n_I = self.kwargs.get('n_I', 0)
            self.p_val = self.kwargs.get('p_val', 1)
            self.isAbove = self.kwargs.get('isAbove', True)
        else:
            self.Cr = self.kwargs.get('Cr', 0)
            self.Cvert = self.kwargs.get('Cvert', 0)
            self.Cph = self.kwargs.get('Cph', 0)
            if self.Cr == self.Cvert == self.Cph == 0.:
                raise Exception("all field coefficients are 0!")

    def table(self):
        """device table of the 'fromfile' and 'bz_monopole' models (None for the closed forms)"""
        if self.fieldtype not in ('fromfile', 'bz_monopole'):
            return None
        if self._table is None:
            from .models import device_table
            if self.fieldtype == 'bz_monopole':
                rs, fs = self.fr_table if self.fr_table is not None else bz_fr_table()
                self._table = device_table((np.asarray(rs, dtype=np.float64), np.asarray(fs, dtype=np.float64)))
            else:
                c = self.cached_data
                self._table = device_table((c['radii'], c['Br'], c['Bth'], c['Bph']))
        return self._table

    def descriptor(self):
        f = self.fieldtype
        if f == 'fromfile':
            return _lib.BF['fromfile'], []
        if f == 'bz_monopole':
            return _lib.BF['bz_monopole'], [float(self.C), float(bool(self.secondorder_only))]
        if f == 'bz_para':
            return _lib.BF['bz_para'], [float(self.C)]
        if f == 'power':
            return _lib.BF['power'], [float(self.pval), float(self.C), float(bool(self.usemono))]
        if f in ('rad', 'vert', 'tor', 'simple'):
            return _lib.BF['simple'], [float(self.Cr), float(self.Cvert), float(self.Cph)]
        if f == 'simple_rm1':
            return _lib.BF['simple_rm1'], [float(self.Cr), float(self.Cvert), float(self.Cph)]
        if f == 'bz_guess':
            return _lib.BF['bz_guess'], [float(self.C)]
        if f == 'gen_power':
            return _lib.BF['gen_power'], [float(self.n_I), float(self.p_val), float(bool(self.isAbove))]
        return _lib.BF['monopoleA'], [float(self.C), float(self.omegafac)]

    def _run(self, a, r, velocity, want_bsq):
        from .models import pack_models
        if not isinstance(r, np.ndarray):
            r = np.array([r]).flatten()
        L = _lib.lib()
        M = pack_models(velocity=velocity, bfield=self)
        dev = _engine.device()
        dr = _engine.to_device(r.ravel(), dev)
        n = int(dr.numel())
        outs = [_engine.empty((n,), dev) for _ in range(3)]
        bsq = _engine.empty((n,), dev) if want_bsq else None
        _lib.check(L.kgeo_bfield_lab(ctypes.byref(M), float(a), n, _engine.dptr(dr), *[_engine.dptr(o) for o in outs],
                                     _engine.dptr(bsq), _engine.stream_ptr()))
        if want_bsq:
            return bsq.cpu().numpy().reshape(r.shape)
        return tuple(o.cpu().numpy().reshape(r.shape) for o in outs)

    def bfield_lab(self, a, r, th=np.pi / 2):
        """lab frame b field starF^i0 = (B1, B2, B3)"""
        _checks.require_equatorial(th, "Bfield.bfield_lab")
        return self._run(a, r, None, False)

    def omega_field(self, a, r, th=np.pi / 2):
        """fieldline angular speed (kgeo/bfields.py:175-197)"""
        if self.fieldtype not in ('bz_monopole', 'monopoleA', 'bz_guess', 'bz_para', 'power', 'gen_power'):
            raise Exception("self.omega_field not implemented for fieldtype %s'!" % self.fieldtype)
        _checks.require_equatorial(th, "Bfield.omega_field")
        from .models import pack_models
        if not isinstance(r, np.ndarray):
            r = np.array([r]).flatten()
        L = _lib.lib()
        M = pack_models(bfield=self)
        dev = _engine.device()
        dr = _engine.to_device(r.ravel(), dev)
        n = int(dr.numel())
        out = _engine.empty((n,), dev)
        _lib.check(L.kgeo_bfield_omega(ctypes.byref(M), float(a), n, _engine.dptr(dr), _engine.dptr(out), _engine.stream_ptr()))
        return out.cpu().numpy().reshape(r.shape)

    def efield_lab(self, a, r, th=np.pi / 2):
        """lab frame electric field F^{0i} of a stationary, axisymmetric field (kgeo/bfields.py:199-235), in the
        equatorial plane: assembled from the device's B^i and omega"""
        _checks.require_equatorial(th, "Bfield.efield_lab")
        if not isinstance(r, np.ndarray):
            r = np.array([r]).flatten()
        if self.fieldtype == 'bz_monopole' and self.secondorder_only:   # Efield_BZmonopole, kgeo/bfields.py:595-606
            return (0, a * np.sin(th) * (-1. / 8. + 2. / r ** 3) / (r ** 2 - 2 * r), 0)
        omega = self.omega_field(a, r, th=th)
        B1, B2, B3 = self.bfield_lab(a, r, th=th)
        a2, r2 = a ** 2, r ** 2
        cth2, sth2 = np.cos(th) ** 2, np.sin(th) ** 2
        Delta = r2 - 2 * r + a2
        Sigma = r2 + a2 * cth2
        Pi = (r2 + a2) ** 2 - a2 * Delta * sth2
        omegaz = 2 * a * r / Pi
        E1 = (omega - omegaz) * Pi * np.sin(th) * B2 / Sigma
        E2 = -(omega - omegaz) * Pi * np.sin(th) * B1 / (Sigma * Delta)
        return (E1, E2, np.zeros_like(E2))

    def bsq(self, a, r, velocity, th=np.pi / 2.):
        """b^2 in the frame u^mu of `velocity`, ideal MHD"""
        _checks.require_equatorial(th, "Bfield.bsq")
        return self._run(a, r, velocity, True)


# ---- the reference's module-level field configurations of the image path (kgeo/bfields.py:429-542, 712-746): same signatures,
# evaluated by the device library in the equatorial plane (th must be pi/2, as everywhere in this package) ----
def _spin_check(a):
    if not (isinstance(a, float) and (0 <= np.abs(a) < 1)):
        raise Exception("|a| should be a float in range [0,1)")


def Bfield_simple(a, r, coeffs):
    """magnetic field vector (Br, Bth, Bph) in the lab frame in equatorial plane, simple configurations: |Br| falls off as
    1/r^2, |Bvert| and |Bph| as 1/r; coeffs = (arad, avert, ator)"""
    _spin_check(a)
    (arad, avert, ator) = coeffs
    return Bfield('simple', Cr=arad, Cvert=avert, Cph=ator).bfield_lab(a, r)


def Bfield_simple_rm1(a, r, coeffs):
    """as Bfield_simple, all components fall off as 1/r"""
    _spin_check(a)
    (arad, avert, ator) = coeffs
    return Bfield('simple_rm1', Cr=arad, Cvert=avert, Cph=ator).bfield_lab(a, r)


def Bfield_BZmagic(a, r, th, C=1):
    """guess ratio for Bphi/Br from the split monopole: (Br, Bth, Bph, OmegaBZ)"""
    _spin_check(a)
    bf = Bfield('bz_guess', C=C)
    return bf.bfield_lab(a, r, th) + (bf.omega_field(a, r, th),)


def Bfield_gen_power(a, r, th, n_I, p_val, isAbove):
    """generalised power-law stream function: (Br, Bth, Bph, Omega_H / 2)"""
    _spin_check(a)
    bf = Bfield('gen_power', n_I=n_I, p_val=p_val, isAbove=isAbove)
    return bf.bfield_lab(a, r, th) + (bf.omega_field(a, r, th),)
