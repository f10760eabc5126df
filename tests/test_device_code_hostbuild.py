"""The device code on the host: the kernels' per-ray and per-sample / per-pixel functions are __host__ __device__, so the same
headers are compiled with g++ (tests/hostsim/*.cpp, no FMA contraction) and compared, WITHOUT a GPU, with fixtures of the unmodified
reference under the gates of the GPU tests:
  * off-equatorial jet physics (kg_offeq.cuh: metric, BZ monopole / power-law field with its field-line speed, drift frame, redshift,
    tetrad, Penrose-Walker transport, per-crossing Stokes parameters) on the crossings of tests/golden/jet_pipeline.npz;
  * the track kernels' setup and per-sample code (theta-function path and Carlson path) on tests/golden/tracks_*.npz;
  * the lensing / image kernels' per-pixel code (geometry and emission, every closed-form model) on lensing.npz / image_*.npz;
  * the two-launch setup == the one-pass setup, bit for bit.
The GPU side of the same comparisons is tests/test_gpu_*.py; the product never loads these host builds."""
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


@pytest.fixture(scope="module")
def hostsim_lib(tmp_path_factory):
    """tests/hostsim/hostsim.cpp (the per-ray / per-sample device functions of the track kernels) compiled once with g++."""
    import ctypes
    so = str(tmp_path_factory.mktemp("hostsim") / "_hostsim.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "kgeo_b200", "csrc"),
                    os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"), "-o", so], check=True, capture_output=True, timeout=600)
    return ctypes.CDLL(so)


def test_split_ray_setup_equals_the_whole_on_the_host(hostsim_lib):
    """The track setup runs as two launches (conserved quantities + polar motion | radial motion).  Host build of the same
    functions: both parts into one workspace == the one-pass setup, every slot and flag bit for bit, for all root classes,
    vortical rays, zero spin and the southern observer."""
    import ctypes
    import numpy as np
    L = hostsim_lib
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


@pytest.mark.parametrize("name,force", [("grid17", 0), ("grid17", 1), ("extremal", 0), ("classII", 0), ("spin0", 0), ("south", 0)])
def test_device_track_code_on_the_host_against_reference(golden, hostsim_lib, name, force):
    """The per-ray setup and per-sample code of the track kernels are __host__ __device__: compiled with g++ (no FMA contraction)
    and run over the rays of a fixture of the UNMODIFIED reference, they meet the gates of the GPU tests -- the kernels' arithmetic
    checked without a GPU.  force = 1 clears the theta-function flags (the Carlson duplication path for every sample)."""
    import ctypes
    import numpy as np
    from helpers import assert_tracks_close
    L = hostsim_lib
    P = ctypes.POINTER(ctypes.c_double)
    g = golden('tracks_' + name)
    a, th_o, r_o, ngeo = [float(x) for x in g['par']]
    A, B = np.ascontiguousarray(g['alpha'], dtype=float), np.ascontiguousarray(g['beta'], dtype=float)
    n, ngeo = A.size, int(ngeo)
    eta = (A ** 2 - a ** 2) * np.cos(th_o) ** 2 + B ** 2
    mode = (1 if np.any(eta > 0) else 0) | (2 if np.any(eta <= 0) else 0)
    if abs(a) < 1e-4 and np.any(eta < 0):
        mode = 1 | 4                                 # the low-spin clamp of raytrace_ana (kerr_raytracing_ana.py:80-82)
    out = {k: np.zeros((ngeo, n)) for k in ('t', 'r', 'th', 'ph', 'sig')}
    rc = L.hostsim_tracks(ctypes.c_double(a), ctypes.c_double(r_o), ctypes.c_double(th_o), ctypes.c_size_t(n), A.ctypes.data_as(P),
                          B.ctypes.data_as(P), mode, ngeo, ctypes.c_double(1. - 1.e-10), force,
                          *[out[k].ctypes.data_as(P) for k in ('t', 'r', 'th', 'ph', 'sig')], None, None)
    assert rc == 0
    ref = {k: g[k] for k in ('tau', 'sig', 't', 'r', 'th', 'ph')}
    ref.update(a=a, r_o=r_o)
    out['tau'] = g['tau']                            # the samples are linspace(0, maxtaufrac * tau_tot, ngeo) in both
    assert_tracks_close(out, ref, what='host build, %s, force %d' % (name, force), pixels=(th_o, A, B))


@pytest.fixture(scope="module")
def hostsim_image_lib(tmp_path_factory):
    """tests/hostsim/hostsim_image.cpp (device code of the lensing / image kernels) compiled once with g++."""
    import ctypes
    so = str(tmp_path_factory.mktemp("hostsim_image") / "_hostsim_image.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(ROOT, "tests", "hostsim", "hostsim_image.cpp"),
                    "-o", so], check=True, capture_output=True, timeout=600)
    return ctypes.CDLL(so)


def test_device_lensing_code_on_the_host_against_reference(golden, hostsim_image_lib):
    """r_equatorial's device code (per-pixel setup, crossing time, Nmax, crossing radius: the geometry kernel of the image) on the
    host against equatorial_lensing.r_equatorial of the unmodified reference, n = 0..2, with the gates of the GPU test."""
    import ctypes
    import numpy as np
    from helpers import TRACK_RTOL, assert_close
    L = hostsim_image_lib
    P = ctypes.POINTER(ctypes.c_double)
    g = golden('lensing')
    for i in range(int(g['ncases'])):
        a, th_o, r_o = [float(x) for x in g['c%d_par' % i]]
        A, B = np.ascontiguousarray(g['c%d_alpha' % i], dtype=float), np.ascontiguousarray(g['c%d_beta' % i], dtype=float)
        n = A.size
        rs, Ir, N = np.zeros((3, n)), np.zeros((3, n)), np.zeros((3, n))
        Imax, npol = np.zeros(n), np.zeros(n)
        L.hs_requatorial(ctypes.c_double(a), ctypes.c_double(r_o), ctypes.c_double(th_o), n, A.ctypes.data_as(P), B.ctypes.data_as(P), 0, 2,
                         rs.ctypes.data_as(P), Ir.ctypes.data_as(P), N.ctypes.data_as(P), Imax.ctypes.data_as(P), npol.ctypes.data_as(P))
        for mbar in range(3):
            ref = g['c%d_m%d' % (i, mbar)]
            assert np.array_equal(N[mbar], ref[3]), "Nmax differs (case %d, m=%d)" % (i, mbar)
            assert_close(rs[mbar], ref[0], TRACK_RTOL, floor=1e-6, what='r_s %d/%d' % (i, mbar))
            assert_close(Ir[mbar], ref[1], 1e-13, what='Ir %d/%d' % (i, mbar))
            assert_close(Imax, ref[2], 2e-10, what='Imax %d/%d' % (i, mbar))


@pytest.mark.parametrize('name', ['default', 'testpol', 'subkep_pol', 'kep_vert_path', 'thermal_genpower', 'infall_tor', 'gelles_simple',
                                  'simfit_rm1', 'spin0_default'])
def test_device_image_code_on_the_host_against_reference(golden, hostsim_image_lib, name):
    """The image kernels' per-pixel code (geometry + emission: image_order) on the host against make_image of the unmodified reference
    for every closed-form model combination of the fixtures, with the gates of the GPU test."""
    import ctypes
    import numpy as np
    import kgeo_b200 as kg
    from kgeo_b200 import _lib, equatorial_images as ei
    from kgeo_b200.models import pack_models
    from helpers import IMAGE_RTOL, TRACK_RTOL, assert_close, image_models
    L = hostsim_image_lib
    P = ctypes.POINTER(ctypes.c_double)
    g = golden('image_' + name)
    a, r_o, th_o, mbar_max, half, npx = [float(x) for x in g['par']]
    em, vel, bf, kw = image_models(g['spec'])
    pop = lambda d: {k: v for k, v in d.items() if k != 'type'}
    M = pack_models(kg.Emissivity(em['type'], **pop(em)), kg.Velocity(vel['type'], **pop(vel)), kg.Bfield(bf['type'], **pop(bf)))
    opts = _lib.ImageOpts(0, int(mbar_max), int(bool(kw.get('polarization', False))), int(bool(kw.get('pathlength', False))),
                          float(kw.get('specind', ei.SPECIND)), float(kw.get('diskangle', ei.DISKANGLE)), float(kw.get('nu_obs', ei.OBSFREQ)))
    A, B = ei.pixel_grid(-half, half, -half, half, 2 * half / npx)
    n, norder = A.size, int(mbar_max) + 1
    planes = {k: np.zeros((norder, n)) for k in ('I', 'Q', 'U', 'r', 'Ir', 'g', 'sinthb', 'lp', 'Ie')}
    single = {k: np.zeros(n) for k in ('nmax', 'npol', 'imax')}
    L.hs_image(ctypes.c_double(a), ctypes.c_double(r_o), ctypes.c_double(th_o), ctypes.byref(M), ctypes.byref(opts), n,
               A.ctypes.data_as(P), B.ctypes.data_as(P), *[planes[k].ctypes.data_as(P) for k in ('I', 'Q', 'U', 'r', 'Ir', 'g', 'sinthb')],
               single['nmax'].ctypes.data_as(P), single['npol'].ctypes.data_as(P), single['imax'].ctypes.data_as(P),
               planes['lp'].ctypes.data_as(P), planes['Ie'].ctypes.data_as(P))
    out = [planes[k].T for k in ('I', 'Q', 'U', 'r', 'Ir', 'g', 'sinthb')] + [single['nmax'], single['npol']]
    names = ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np']
    if kw.get('pathlength'):
        out += [planes['lp'].T, planes['Ie'].T]
        names += ['lp', 'Ie']
    ref_I = g['I']
    for got, key in zip(out, names):
        ref = g[key]
        assert got.shape == ref.shape, key
        assert np.array_equal(got == 0, ref == 0), "%s %s: zero mask differs" % (name, key)
        if key in ('Q', 'U'):
            den = np.where(ref_I != 0, np.abs(ref_I), 1.0)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), key
            assert np.nanmax(np.abs(got - ref) / den) <= IMAGE_RTOL, key
        elif key == 'n':
            assert np.array_equal(got, ref)
        else:
            assert_close(got, ref, IMAGE_RTOL if key in ('I', 'g', 'sinthb', 'lp', 'Ie', 'np') else TRACK_RTOL, floor=1e-300,
                         what='%s %s' % (name, key))


def test_device_model_code_on_the_host_against_reference(golden, hostsim_image_lib):
    """The velocity and emissivity closed forms of the image kernels (kg_image.cuh) on the host against the reference's module-level
    u_zamo ... u_general (velocities.py:179-470) and emisBPL / emisGLM (emissivities.py:121-139)."""
    import ctypes
    import numpy as np
    import kgeo_b200 as kg
    from kgeo_b200.models import pack_models
    from helpers import assert_close
    L = hostsim_image_lib
    P = ctypes.POINTER(ctypes.c_double)
    g = golden('model_functions')
    null = ctypes.cast(None, P)
    for a in (0.0, 0.5, 0.94):
        tag = 'a%02d' % int(round(100 * a))
        r = np.ascontiguousarray(g[tag + '_r'], dtype=float)
        n = r.size
        vels = dict(zamo=kg.Velocity('zamo'), infall=kg.Velocity('infall'), kep=kg.Velocity('kep'), kep_retro=kg.Velocity('kep', retrograde=True),
                    subkep=kg.Velocity('subkep', fac_subkep=0.7), gelles=kg.Velocity('gelles', gelles_beta=0.4, gelles_chi=-2.0),
                    simfit=kg.Velocity('simfit'), general=kg.Velocity('general', fac_subkep=0.7, beta_phi=0.6, beta_r=0.8))
        for k, v in vels.items():
            M = pack_models(velocity=v)
            u = np.zeros((4, n))
            L.hs_velocity(ctypes.byref(M), ctypes.c_double(a), n, r.ctypes.data_as(P), u.ctypes.data_as(P))
            assert_close(u, g[tag + '_u_' + k], 1e-10, floor=1e-12, what='%s %s' % (tag, k))
        emis = dict(bpl=kg.Emissivity('bpl'), bpl2=kg.Emissivity('bpl', p1=-1.5, p2=-0.3), glm=kg.Emissivity('glm'),
                    ring=kg.Emissivity('ring', r_ring=4.5, gamma_off=0.2, sigma=0.7))
        for k, e in emis.items():
            M = pack_models(emissivity=e)
            j = np.zeros(n)
            L.hs_emis(ctypes.byref(M), ctypes.c_double(a), n, r.ctypes.data_as(P), null, null, null, ctypes.c_double(230e9),
                      ctypes.c_double(1.0), j.ctypes.data_as(P))
            assert_close(j, g[tag + '_' + k], 1e-10, floor=1e-300, what='%s %s' % (tag, k))


def test_device_bfield_code_on_the_host_against_reference(golden, hostsim_image_lib):
    """The B-field closed forms of the image kernels on the host against the reference's module-level Bfield_simple, Bfield_simple_rm1,
    Bfield_BZmagic and Bfield_gen_power (bfields.py:429-542, 712-746), field-line angular speed included."""
    import ctypes
    import numpy as np
    import kgeo_b200 as kg
    from kgeo_b200.models import pack_models
    from helpers import assert_close
    L = hostsim_image_lib
    P = ctypes.POINTER(ctypes.c_double)
    g = golden('model_functions')
    for a in (0.0, 0.5, 0.94):
        tag = 'a%02d' % int(round(100 * a))
        r = np.ascontiguousarray(g[tag + '_r'], dtype=float)
        n = r.size
        fields = dict(simple=kg.Bfield('simple', Cr=0.87, Cvert=0.3, Cph=-0.5), simple_rm1=kg.Bfield('simple_rm1', Cr=0.4, Cvert=-0.7, Cph=0.2))
        if a > 0:
            fields.update(bzmagic=kg.Bfield('bz_guess', C=-1.5), gen_power=kg.Bfield('gen_power', n_I=0.5, p_val=0.8, isAbove=True))
        for k, bf in fields.items():
            M = pack_models(bfield=bf)
            B, om = np.zeros((3, n)), np.zeros(n)
            L.hs_bfield_omega(ctypes.byref(M), ctypes.c_double(a), n, r.ctypes.data_as(P), B.ctypes.data_as(P), om.ctypes.data_as(P))
            ref = g[tag + '_B_' + k]
            assert_close(B, ref[:3], 1e-10, floor=1e-6 * np.max(np.abs(ref[:3])), what='%s B %s' % (tag, k))
            if ref.shape[0] == 4:
                assert_close(om, ref[3], 1e-12, what='%s omega %s' % (tag, k))


def test_device_special_functions_on_the_host_against_scipy_fixture(golden, tmp_path):
    """The special-function layer of the kernels (kg_special.cuh: Carlson RF/RD/RJ/RC, Legendre F/E/K, the reference's Pi with its
    special-value rules, Jacobi sn/cn/dn/am) on the host against the values the reference's SciPy calls produce
    (tests/golden/special.npz), with the gates of tests/test_gpu_special.py."""
    import ctypes
    import numpy as np
    from helpers import rel_err
    so = str(tmp_path / "_hostsim_special.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "kgeo_b200", "csrc"),
                    os.path.join(ROOT, "tests", "hostsim", "hostsim_special.cpp"), "-o", so], check=True, capture_output=True, timeout=600)
    L = ctypes.CDLL(so)
    P = ctypes.POINTER(ctypes.c_double)
    g = golden('special')

    def call(fn, *arrs, nout=1):
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=float), np.shape(arrs[0])), dtype=float) for x in arrs]
        outs = [np.zeros(arrs[0].shape) for _ in range(nout)]
        getattr(L, fn)(arrs[0].size, *[x.ctypes.data_as(P) for x in arrs], *[o.ctypes.data_as(P) for o in outs])
        return outs[0] if nout == 1 else outs

    assert rel_err(call('hs_rf', g['c_x'], g['c_y'], g['c_z']), g['rf']) < 4e-15
    assert rel_err(call('hs_rd', g['c_x'], g['c_y'], g['c_z']), g['rd']) < 4e-15
    assert rel_err(call('hs_rj', g['c_x'], g['c_y'], g['c_z'], g['c_p']), g['rj']) < 2e-14
    assert rel_err(call('hs_rc', g['c_x'], g['c_y']), g['rc']) < 4e-15
    assert rel_err(call('hs_F', g['fe_phi'], g['fe_m']), g['F']) < 1e-14
    assert rel_err(call('hs_E', g['fe_phi'], g['fe_m']), g['E'], floor=1e-2) < 1e-14
    assert rel_err(call('hs_K', g['fe_m']), g['K']) < 4e-15
    assert rel_err(call('hs_Pi', np.full(18, .8), g['pi_list_phi'], np.full(18, -.33)), g['pi_list']) < 2e-15
    assert rel_err(call('hs_Pi', g['pi_n'], g['pi_phi'], g['pi_m']), g['pi'], floor=0.1) < 2e-13
    got, ref = call('hs_Pi', g['pisp_n'], g['pisp_phi'], g['pisp_m']), g['pisp']
    keep = np.array([0, 1, 2, 4, 5, 6, 7])           # entry 3: n == 1 at |phi| < pi/2, ill-conditioned in the reference itself
    assert rel_err(got[keep], ref[keep]) < 1e-13
    sn, cn, dn, am = call('hs_jacobi', g['j_u'], g['j_m'], nout=4)
    assert np.max(np.abs(sn - g['sn'])) < 2e-14 and np.max(np.abs(cn - g['cn'])) < 2e-14
    assert np.max(np.abs(dn - g['dn'])) < 2e-12       # SciPy's own dn loses digits as m -> 1
    assert np.max(np.abs(am - g['am'])) < 3e-14


def test_device_crossing_counts_on_the_host_against_reference(golden, hostsim_image_lib):
    """n_equatorial_crossings / n_poloidal_orbits device code (kerr_raytracing_utils.py:551-684) on the host, full (ntau, npix) tau."""
    import ctypes
    import numpy as np
    from helpers import assert_close
    L = hostsim_image_lib
    P = ctypes.POINTER(ctypes.c_double)
    g = golden('counts')
    a, th_o = [float(x) for x in g['par']]
    A, B = np.ascontiguousarray(g['alpha'], dtype=float), np.ascontiguousarray(g['beta'], dtype=float)
    tau = np.ascontiguousarray(g['tau'], dtype=float)
    neq, npol = np.zeros(tau.shape), np.zeros(tau.shape)
    L.hs_counts(ctypes.c_double(a), ctypes.c_double(th_o), A.size, tau.shape[0], A.ctypes.data_as(P), B.ctypes.data_as(P),
                tau.ctypes.data_as(P), neq.ctypes.data_as(P), npol.ctypes.data_as(P))
    assert np.array_equal(neq, g['neq'])
    assert_close(npol, g['npol'], 1e-12, what='n_poloidal')
