"""CPU check of the off-equatorial jet physics written for the rest of SURVEY 8 f4 (kgeo_b200/csrc/kg_offeq.cuh: metric, BZ monopole /
power-law field with its field-line speed, drift frame, redshift, tetrad, Penrose-Walker transport, per-crossing Stokes parameters).
The header's functions are __host__ __device__: it is compiled for the host with g++ and compared with the pinned oracle on the
crossings of tests/golden/jet_pipeline.npz (and with the unmodified reference's raw Stokes parameters stored there).  The GPU side of
the same comparison is tests/test_gpu_jet.py."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_offeq_device_functions_on_the_host(tmp_path):
    so = str(tmp_path / "_hostsim_offeq.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "kgeo_b200", "csrc"),
                    os.path.join(ROOT, "tests", "hostsim", "hostsim_offeq.cpp"), "-o", so], check=True, capture_output=True, timeout=300)
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    import check_offeq
    worst = check_offeq.run(so, verbose=False)
    assert worst < 1e-10, worst


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_ff_boost_and_gammamax_on_the_host(tmp_path, golden):
    """The device code of nu_parallel='FF' (getnu_cons, kgeo/velocities.py:474-486) and of the gammamax cap (:593-611), compiled
    for the host and run in the kernels' order (gamma pre-pass, first argmin, density + Stokes parameters), against the raw
    per-crossing Stokes parameters of the UNMODIFIED reference's getstokes (tests/golden/jet_boost.npz)."""
    import ctypes
    import json
    import warnings
    import numpy as np
    so = str(tmp_path / "_hostsim_offeq2.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "kgeo_b200", "csrc"),
                    os.path.join(ROOT, "tests", "hostsim", "hostsim_offeq.cpp"), "-o", so], check=True, capture_output=True, timeout=300)
    from oracle import kgeo_oracle as ko
    from kgeo_b200 import off_eq
    warnings.simplefilter('ignore')
    np.seterr(all='ignore')
    L = ctypes.CDLL(so)
    L.hs_jet_stokes.restype = ctypes.c_int
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda x: x.ctypes.data_as(dp)
    D = ctypes.c_double
    g, gp = golden('jet_boost'), golden('jet_pipeline')
    a, r_o, th_o, psit, neqmax, ngeo = [float(x) for x in g['par']]
    frx, frf = [np.ascontiguousarray(v) for v in ko.bz_fr_table()]
    shape = tuple(int(x) for x in gp['mono_shape'])
    nguess, npix = shape
    A, B = np.ascontiguousarray(g['alpha'].flatten()), np.ascontiguousarray(g['beta'].flatten())
    r, th = np.ascontiguousarray(gp['mono_root_r']), np.ascontiguousarray(gp['mono_root_th'])
    kr, kth = np.ascontiguousarray(gp['mono_root_signpr']), np.ascontiguousarray(gp['mono_root_signpth'])
    field, r0, norm = off_eq.jet_scalars('mono', r, th, shape, psit, a, 2, 1.0)
    for name, kw in zip(json.loads(str(g['names'])), json.loads(str(g['kwargs']))):
        nu = kw['nu_parallel']
        ff = isinstance(nu, str)
        gmax = kw.get('gammamax') or 0.0
        ffg, ffr = np.zeros(3), np.zeros(3 * nguess)
        if ff:
            rf = np.nan_to_num(r)
            i0 = np.where(rf != 0)[0][0]
            ffg = np.array(off_eq.ff_constants_mono(rf[i0], th[i0], a))
            rn, tn = rf.reshape(shape), np.arccos(np.abs(np.cos(th.reshape(shape))))
            ffr = np.ascontiguousarray(np.array([off_eq.ff_constants_mono(rn[m][np.where(rn[m] != 0)[0][0]], tn[m][np.where(rn[m] != 0)[0][0]], a)
                                                 for m in range(nguess)]).ravel())
        outs = [np.empty(r.size) for _ in range(5)]
        bad = L.hs_jet_stokes(0, D(1.0), D(1.0), 0, P(frx), P(frf), ctypes.c_longlong(frx.size), D(a), D(th_o), npix, nguess, P(A), P(B),
                              P(r), P(th), P(kr), P(kth), D(0.0 if ff else nu), int(ff), P(ffg), P(ffr), D(gmax), D(2.0), D(r0), D(norm),
                              D(psit), D(1.0), *[P(o) for o in outs])
        assert bad == 0
        for k, got in zip(('i', 'q', 'u'), outs[:3]):
            ref = g[name + '_raw_' + k]
            assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), (name, k)


def test_split_ray_setup_equals_the_whole_on_the_host(tmp_path):
    """The track setup runs as two launches (conserved quantities + polar motion | radial motion).  Host build of the same
    functions: both parts into one workspace == the one-pass setup, every slot and flag bit for bit, for all root classes,
    vortical rays, zero spin and the southern observer."""
    import ctypes
    import numpy as np
    so = str(tmp_path / "_hostsim.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "kgeo_b200", "csrc"),
                    os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"), "-o", so], check=True, capture_output=True, timeout=600)
    L = ctypes.CDLL(so)
    NW = L.hostsim_slots()
    P, I = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    seen = 0
    for a, th_o, half, mode in ((0.94, 17 * np.pi / 180, 8., 3), (0.998, 10 * np.pi / 180, 3., 3), (0.0, 45 * np.pi / 180, 8., 7),
                                (0.7, 135 * np.pi / 180, 8., 3), (0.94, 80 * np.pi / 180, 3., 3)):
        x = np.linspace(-half, half, 29)
        A, B = [v.flatten().copy() for v in np.meshgrid(x, x)]
        n = A.size
        w0, w1 = np.full((n, NW), -123.0), np.full((n, NW), -123.0)
        f0, f1 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        L.hostsim_setup_parts(ctypes.c_double(a), ctypes.c_double(1000.), ctypes.c_double(th_o), ctypes.c_size_t(n),
                              A.ctypes.data_as(P), B.ctypes.data_as(P), mode, w0.ctypes.data_as(P), f0.ctypes.data_as(I),
                              w1.ctypes.data_as(P), f1.ctypes.data_as(I))
        assert np.array_equal(f0, f1)
        assert ((w0 == w1) | (np.isnan(w0) & np.isnan(w1))).all()
        seen |= int(np.bitwise_or.reduce(f0))
    assert seen & 4 and seen & 8 and seen & 16 and seen & 1 and seen & 2      # radial cases 2, 3, 4; ordinary and vortical polar motion
