"""Velocity model descriptor with the constructor and u_lab interface of kgeo/velocities.py:23-96.
The closed forms, the tabulated 'fromfile' model and the E x B 'driftframe' of a B-field model (numeric nu_parallel) run on
the device (kg_image.cuh)."""
import ctypes

import numpy as np

from . import _checks, _engine, _lib

# defaults of the reference (kgeo/velocities.py:14-21)
BETA = 0.3
CHI = -150 * (np.pi / 180.)
ELLISCO = 1.
VRISCO = 2.
P1 = 6.
P2 = 2.
DD = 0.2


class Velocity(object):
    """lab frame velocity as a function of r in the equatorial plane"""

    def __init__(self, veltype="kep", **kwargs):
        self.veltype = veltype
        self.kwargs = kwargs
        if self.veltype == 'zamo' or self.veltype == 'infall':
            pass
        elif self.veltype == 'kep' or self.veltype == 'cunningham':
            self.retrograde = self.kwargs.get('retrograde', False)
        elif self.veltype == 'subkep' or self.veltype == 'cunningham_subkep':
            self.retrograde = self.kwargs.get('retrograde', False)
            self.fac_subkep = self.kwargs.get('fac_subkep', 1)
        elif self.veltype == 'general':
            self.retrograde = self.kwargs.get('retrograde', False)
            self.fac_subkep = self.kwargs.get('fac_subkep', 1)
            self.beta_phi = self.kwargs.get('beta_phi', 1)
            self.beta_r = self.kwargs.get('beta_r', 1)
        elif self.veltype == 'gelles':
            self.gelles_beta = self.kwargs.get('gelles_beta', BETA)
            self.gelles_chi = self.kwargs.get('gelles_chi', CHI)
        elif self.veltype == 'simfit':
            self.ell_isco = self.kwargs.get('ell_isco', ELLISCO)
            self.vr_isco = self.kwargs.get('vr_isco', VRISCO)
            self.p1 = self.kwargs.get('p1', P1)
            self.p2 = self.kwargs.get('p2', P2)
            self.dd = self.kwargs.get('dd', DD)
        elif self.veltype == 'fromfile':
            from .models import load_cache_from_file
            self.filename = self.kwargs.get('file', None)
            radii, U1, U2, U3 = load_cache_from_file(self.filename, ('U1', 'U2', 'U3'))
            self.cached_data = dict(radii=radii, U1=U1, U2=U2, U3=U3)
            self._table = None
        elif self.veltype == 'driftframe':
            from .bfields import Bfield
            self.bfield = self.kwargs.get('bfield', None)
            if self.bfield is None:
                self.bfield = Bfield('bz_para')                       # BFIELD_DEFAULT, kgeo/velocities.py:14
            self.nu_parallel = self.kwargs.get('nu_parallel', 0)
            self.gammamax = self.kwargs.get('gammamax', None)
            if isinstance(self.nu_parallel, str) or self.gammamax:
                raise NotImplementedError("Velocity('driftframe') with nu_parallel='FF' or a gammamax needs the jet "
                                          "model's stagnation-surface root finds (kgeo/ff_boost.py; SURVEY 8 f4)")
            if not -1 <= self.nu_parallel <= 1:
                raise Exception("nu_parallel should be in the range (-1,1)")
        else:
            raise Exception("veltype %s not recognized in Velocity!" % self.veltype)

    def table(self):
        """device table of the 'fromfile' model (None for the closed forms)"""
        if self.veltype != 'fromfile':
            return None
        if self._table is None:
            from .models import device_table
            c = self.cached_data
            self._table = device_table((c['radii'], c['U1'], c['U2'], c['U3']))
        return self._table

    def descriptor(self):
        """(type id, parameter list) for kgeo_models.velocity_*"""
        v = self.veltype
        if v == 'fromfile':
            return _lib.VEL['fromfile'], []
        if v == 'driftframe':
            return _lib.VEL['driftframe'], [float(self.nu_parallel)]
        if v == 'zamo':
            return _lib.VEL['zamo'], []
        if v == 'infall':
            return _lib.VEL['infall'], []
        if v in ('kep', 'cunningham'):
            return _lib.VEL['kep'], [float(bool(self.retrograde))]
        if v in ('subkep', 'cunningham_subkep'):
            return _lib.VEL['subkep'], [float(bool(self.retrograde)), float(self.fac_subkep)]
        if v == 'general':
            return _lib.VEL['general'], [float(bool(self.retrograde)), float(self.fac_subkep), float(self.beta_phi),
                                         float(self.beta_r)]
        if v == 'gelles':
            return _lib.VEL['gelles'], [float(self.gelles_beta), float(self.gelles_chi)]
        return _lib.VEL['simfit'], [float(self.ell_isco), float(self.vr_isco), float(self.p1), float(self.p2),
                                    float(self.dd)]

    def u_lab(self, a, r, th=np.pi / 2., retqty=False):
        """Return lab frame contravariant velocity vector (u0, u1, u2, u3) at radii r (equatorial plane)"""
        from .models import pack_models
        _checks.require_equatorial(th, "Velocity.u_lab")
        if retqty:
            raise NotImplementedError("Velocity.u_lab(retqty=True) is not built (the reference returns intermediate "
                                      "quantities of the drift-frame construction there)")
        if not isinstance(r, np.ndarray):
            r = np.array([r]).flatten()
        L = _lib.lib()
        M = pack_models(velocity=self, bfield=getattr(self, 'bfield', None))
        dev = _engine.device()
        dr = _engine.to_device(r.ravel(), dev)
        n = int(dr.numel())
        outs = [_engine.empty((n,), dev) for _ in range(4)]
        _lib.check(L.kgeo_velocity_u_lab(ctypes.byref(M), float(a), n, _engine.dptr(dr), *[_engine.dptr(o) for o in outs],
                                         _engine.stream_ptr()))
        return tuple(o.cpu().numpy().reshape(r.shape) for o in outs)


# ---- the reference's module-level closed forms (kgeo/velocities.py:179-470): same signatures, evaluated by the device library ----
def _spin_check(a):
    if not (isinstance(a, float) and (0 <= np.abs(a) < 1)):
        raise Exception("|a| should be a float in range [0,1)")


def u_zamo(a, r):
    """velocity for zero angular momentum frame (kgeo/velocities.py:179-208); components are arrays of r's shape"""
    _spin_check(a)
    return Velocity('zamo').u_lab(a, r)


def u_infall(a, r):
    """velocity for geodesic equatorial infall from infinity (:211-226)"""
    _spin_check(a)
    return Velocity('infall').u_lab(a, r)


def u_kep(a, r, retrograde=False):
    """Cunningham velocity for material on keplerian orbits and infalling inside isco (:229-281)"""
    _spin_check(a)
    return Velocity('kep', retrograde=retrograde).u_lab(a, r)


def u_subkep(a, r, fac_subkep=1, retrograde=False):
    """(equatorial) subkeplerian velocities (:284-359)"""
    _spin_check(a)
    return Velocity('subkep', retrograde=retrograde, fac_subkep=fac_subkep).u_lab(a, r)


def u_gelles(a, r, beta=0.3, chi=-150 * (np.pi / 180.)):
    """velocity prescription from Gelles+2021, Eq A4 (:362-389)"""
    _spin_check(a)
    return Velocity('gelles', gelles_beta=beta, gelles_chi=chi).u_lab(a, r)


def u_grmhd_fit(a, r, ell_isco=ELLISCO, vr_isco=VRISCO, p1=P1, p2=P2, dd=DD):
    """velocity for power laws fit to grmhd ell, conserved inside isco (:392-439)"""
    _spin_check(a)
    return Velocity('simfit', ell_isco=ell_isco, vr_isco=vr_isco, p1=p1, p2=p2, dd=dd).u_lab(a, r)


def u_general(a, r, fac_subkep=1, beta_phi=1, beta_r=1, retrograde=False):
    """general velocity model from AART paper, keplerian by default (:442-470)"""
    _spin_check(a)
    return Velocity('general', retrograde=retrograde, fac_subkep=fac_subkep, beta_phi=beta_phi, beta_r=beta_r).u_lab(a, r)
