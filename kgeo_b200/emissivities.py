"""Emissivity model descriptor with the constructor / jrest interface of kgeo/emissivities.py:37-118."""
import ctypes

import numpy as np

from . import _engine, _lib

# defaults of the reference (kgeo/emissivities.py:16-35)
P1E_230 = -2.0
P2E_230 = -0.5
P1E_86 = 0
P2E_86 = -.75
GAMMAOFF = -1.5
SIGMA_GLM = 0.5
R_RING = 4.5
SIGMA_RING = 0.3
OBSFREQ = 230.e9
SPECIND = 0.0


class Emissivity(object):
    """rest frame emissivity as a function of r in the equatorial plane"""

    def __init__(self, emistype="bpl", **kwargs):
        self.emistype = emistype
        self.kwargs = kwargs
        self.emiscut_in = self.kwargs.get('emiscut_in', 0)
        self.emiscut_out = self.kwargs.get('emiscut_out', 1.e10)
        if self.emistype == 'thermal':
            self.Rb = float(self.kwargs.get('Rb', 5.0))
            self.ne0 = float(self.kwargs['ne0'])
            self.Te0 = float(self.kwargs['Te0'])
            self.B0 = float(self.kwargs['B0'])
            self.alpha_n = float(self.kwargs.get('alpha_n', 0.7))
            self.alpha_T = float(self.kwargs.get('alpha_T', 1.0))
            self.alpha_B = float(self.kwargs.get('alpha_B', 1.5))
            self.bfield_model = bool(self.kwargs.get('bfield_model', False))
        elif self.emistype == 'bpl':
            self.p1 = self.kwargs.get('p1', P1E_230)
            self.p2 = self.kwargs.get('p2', P2E_230)
        elif self.emistype == 'ring':
            self.mu_ring = self.kwargs.get('r_ring', R_RING)
            self.gamma_off = self.kwargs.get('gamma_off', 0)
            self.sigma = self.kwargs.get('sigma', SIGMA_RING)
        elif self.emistype == 'glm':
            self.mu_ring = False
            self.gamma_off = self.kwargs.get('gamma_off', GAMMAOFF)
            self.sigma = self.kwargs.get('sigma', SIGMA_GLM)
        elif self.emistype == 'constant':
            pass
        else:
            raise Exception("emistype %s not recognized in Emissivity!" % self.emistype)

    def descriptor(self):
        e = self.emistype
        if e == 'thermal':
            return _lib.EM['thermal'], [self.Rb, self.ne0, self.Te0, self.B0, self.alpha_n, self.alpha_T, self.alpha_B,
                                        float(self.bfield_model)]
        if e == 'bpl':
            return _lib.EM['bpl'], [float(self.p1), float(self.p2)]
        if e == 'ring':
            return _lib.EM['ring'], [float(self.mu_ring), float(self.gamma_off), float(self.sigma)]
        if e == 'glm':
            return _lib.EM['glm'], [float(self.gamma_off), float(self.sigma)]
        return _lib.EM['constant'], []

    def jrest(self, a, r, g=None, sinthetab=None, Bmag=None, nu_obs=OBSFREQ, specind=SPECIND):
        """rest-frame emissivity at radii r; g / sinthetab add the spectral factors as in the reference"""
        from .models import pack_models
        r = np.asarray(r, dtype=np.float64)
        L = _lib.lib()
        M = pack_models(emissivity=self)
        dev = _engine.device()
        dr = _engine.to_device(r.ravel(), dev)
        n = int(dr.numel())

        def opt(x):
            if x is None:
                return None
            return _engine.to_device(np.broadcast_to(np.asarray(x, dtype=np.float64), r.shape).ravel(), dev)

        dg, ds, db = opt(g), opt(sinthetab), opt(Bmag)
        out = _engine.empty((n,), dev)
        _lib.check(L.kgeo_emissivity_jrest(ctypes.byref(M), float(a), n, _engine.dptr(dr), _engine.dptr(dg),
                                           _engine.dptr(ds), _engine.dptr(db), float(nu_obs), float(specind),
                                           _engine.dptr(out), _engine.stream_ptr()))
        return out.cpu().numpy().reshape(r.shape)


# ---- the reference's module-level emissivity profiles (kgeo/emissivities.py:121-139), evaluated by the device library ----
def emisBPL(a, r, p1=P1E_230, p2=P2E_230):
    """emissivity at radius r - broken power law model fit to GRMHD"""
    return Emissivity('bpl', p1=p1, p2=p2).jrest(a, r)


def emisGLM(a, r, gamma_off=GAMMAOFF, sigma=SIGMA_GLM, mu_ring=False):
    """emissivity at radius r from GLM paper (mu_ring: the ring radius instead of r_-)"""
    if mu_ring:
        return Emissivity('ring', r_ring=mu_ring, gamma_off=gamma_off, sigma=sigma).jrest(a, r)
    return Emissivity('glm', gamma_off=gamma_off, sigma=sigma).jrest(a, r)
