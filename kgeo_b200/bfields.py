"""Magnetic field model descriptor with the constructor / bfield_lab / bsq interface of kgeo/bfields.py:33-131,365-396.
Closed-form models and the tabulated 'fromfile' model run on the device; 'bz_monopole', 'bz_para', 'power',
'const_comoving' are outside the accelerated path (a data file the reference does not ship / Lambert W)."""
import ctypes

import numpy as np

from . import _engine, _lib


class Bfield(object):
    """b-field as a function of r in the equatorial plane"""

    def __init__(self, fieldtype="rad", **kwargs):
        self.fieldtype = fieldtype
        if self.fieldtype in ['rad', 'vert', 'tor', 'simple', 'simple_rm1', 'monopoleA', 'bz_guess', 'gen_power', 'fromfile']:
            self.fieldframe = 'lab'
        elif self.fieldtype in ['bz_monopole', 'bz_para', 'power', 'const_comoving']:
            raise NotImplementedError("Bfield(%r) is outside the accelerated path (SURVEY 8f)" % self.fieldtype)
        else:
            raise Exception("fieldtype %s not recognized in Bfield!" % self.fieldtype)
        self.kwargs = kwargs

        if self.fieldtype in ['bz_guess']:
            self.secondorder_only = self.kwargs.get('secondorder_only', False)
            self.C = self.kwargs.get('C', 1)
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
        """device table of the 'fromfile' model (None for the closed forms)"""
        if self.fieldtype != 'fromfile':
            return None
        if self._table is None:
            from .models import device_table
            c = self.cached_data
            self._table = device_table((c['radii'], c['Br'], c['Bth'], c['Bph']))
        return self._table

    def descriptor(self):
        f = self.fieldtype
        if f == 'fromfile':
            return _lib.BF['fromfile'], []
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
        return self._run(a, r, None, False)

    def bsq(self, a, r, velocity, th=np.pi / 2.):
        """b^2 in the frame u^mu of `velocity`, ideal MHD"""
        return self._run(a, r, velocity, True)
