"""CPU-only tests: the C-ABI library loads and exports every symbol include/kgeo_b200.h declares (no compute calls),
the ctypes structs mirror the header, and the host-side logic (argument checks, pixel grid, model packing, setup mode,
row sharding) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "kgeo_b200.h")


@pytest.fixture(scope="module")
def lib():
    from kgeo_b200 import build, _lib
    build.build()            # nvcc cross-compiles for sm_100a without a GPU
    return _lib


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kgeo_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_functions()
    assert len(names) >= 25
    L = ctypes.CDLL(lib.library_path())
    for n in names:
        assert hasattr(L, n), "libkgeo_b200.so does not export %s" % n
        assert n in lib.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(lib.SIGNATURES) == names, "bindings for symbols the header does not declare"


def test_abi_version_and_struct_layout(lib):
    L = lib.lib()
    assert L.kgeo_abi_version() == 4
    assert L.kgeo_ray_workspace_slots() > 30
    assert L.kgeo_launch_count() == 0
    assert ctypes.sizeof(lib.Observer) == 24
    assert ctypes.sizeof(lib.Models) == 16 + 3 * 64 + 16 + 32
    assert ctypes.sizeof(lib.ImageOpts) == 16 + 24
    assert ctypes.sizeof(lib.ShardMap) == 40 and ctypes.sizeof(lib.PeerHandle) == 64
    text = open(HEADER).read()
    for enum, table in (("kgeo_velocity_type", lib.VEL), ("kgeo_bfield_type", lib.BF), ("kgeo_emissivity_type", lib.EM)):
        body = re.search(r"enum %s \{(.*?)\}" % enum, text, re.S).group(1)
        order = [x.strip().split("=")[0].strip() for x in body.split(",")]
        for name, idx in table.items():
            assert order[idx].lower().replace("_", "").endswith(name.lower().replace("_", "")), (enum, name, order[idx])


def test_argument_errors_without_a_gpu(lib):
    """Bad arguments are rejected before any CUDA call, with the reference's messages."""
    L = lib.lib()
    null = ctypes.cast(None, lib.c_double_p)
    obs = lib.Observer(1.5, 1000., 0.3)
    rc = L.kgeo_ray_setup(ctypes.byref(obs), 4, null, null, 3, null, ctypes.cast(None, lib.c_int32_p), None)
    assert rc == lib.KGEO_E_RANGE and b"range (-1,1)" in L.kgeo_last_error()
    obs = lib.Observer(0.5, 10., 0.3)
    assert L.kgeo_ray_setup(ctypes.byref(obs), 4, null, null, 3, null, ctypes.cast(None, lib.c_int32_p), None) == lib.KGEO_E_RANGE
    obs = lib.Observer(0.5, 1000., np.pi / 2)
    assert L.kgeo_ray_setup(ctypes.byref(obs), 4, null, null, 3, null, ctypes.cast(None, lib.c_int32_p), None) == lib.KGEO_E_RANGE
    obs = lib.Observer(0.5, 1000., 0.3)
    assert L.kgeo_ray_setup(ctypes.byref(obs), 4, null, null, 3, null, ctypes.cast(None, lib.c_int32_p), None) == lib.KGEO_E_BADARG
    M = lib.Models()
    M.velocity_type = 99
    assert L.kgeo_velocity_u_lab(ctypes.byref(M), 0.5, 0, null, null, null, null, null, None) == lib.KGEO_E_MODEL
    with pytest.raises(lib.KgeoError, match="unknown velocity model"):
        lib.check(lib.KGEO_E_MODEL)


def test_argument_errors_of_the_round_2_entry_points(lib):
    """The integrals, curve-library, sharded-image and scatter entry points validate before any CUDA call."""
    L = lib.lib()
    null = ctypes.cast(None, lib.c_double_p)
    inull = ctypes.cast(None, lib.c_int32_p)
    x = np.zeros(12)
    px = x.ctypes.data_as(lib.c_double_p)          # never dereferenced: every call below fails validation first
    fl = np.zeros(12, dtype=np.int32).ctypes.data_as(lib.c_int32_p)
    obs = lib.Observer(0.5, 1000., 0.3)
    assert L.kgeo_ray_setup_conserved(ctypes.byref(obs), 4, null, px, null, 3, 1, px, fl, None) == lib.KGEO_E_BADARG
    inf = lib.Observer(0.5, float("inf"), 0.3)
    assert L.kgeo_ray_setup_conserved(ctypes.byref(inf), 4, px, px, null, 3, 1, px, fl, None) == lib.KGEO_E_RANGE
    assert b"finite" in L.kgeo_last_error()
    assert L.kgeo_track_integrals(ctypes.byref(obs), 4, px, fl, 3, null, 4, 1, *([null] * 7), 4, None) == lib.KGEO_E_BADARG
    assert L.kgeo_track_integrals(ctypes.byref(obs), 4, px, fl, 3, px, 4, 1, *([null] * 7), 3, None) == lib.KGEO_E_BADARG
    assert b"leading dimension" in L.kgeo_last_error()
    assert L.kgeo_rho_library(2, px, px, px, 3, 1, 12, px, 2e-12, 1e-15, 100, px, inull, None) == lib.KGEO_E_BADARG
    assert L.kgeo_rho_library(2, px, px, px, 0, 1, 12, px, 0.0, 1e-15, 100, px, inull, None) == lib.KGEO_E_BADARG
    assert L.kgeo_rho_library(0, null, null, null, 0, 1, 12, null, 2e-12, 1e-15, 100, null, inull, None) == 0     # nothing to do
    for smap in (lib.ShardMap(5, 1, 0, 12, 0), lib.ShardMap(4, 0, 0, 12, 0), lib.ShardMap(0, 1, 0, 12, 0)):
        assert L.kgeo_shard_scatter(px, 2, 12, ctypes.byref(smap), px, None) == lib.KGEO_E_BADARG
    assert L.kgeo_shard_scatter(px, 0, 12, ctypes.byref(lib.ShardMap(4, 1, 0, 12, 0)), px, None) == 0
    M = lib.Models()
    opts = lib.ImageOpts(0, 2, 0, 0, 1.0, 0.1, 230e9)
    for smap in (lib.ShardMap(5, 1, 0, 12, 0), lib.ShardMap(4, 0, 0, 12, 0), lib.ShardMap(4, 2, 1, 12, 0)):
        rc = L.kgeo_equatorial_image_sharded(ctypes.byref(inf), ctypes.byref(M), ctypes.byref(opts), 12, px, px, ctypes.byref(smap),
                                             *([null] * 12), None)
        assert rc == lib.KGEO_E_BADARG, L.kgeo_last_error()
    assert L.kgeo_equatorial_image_sharded(ctypes.byref(inf), ctypes.byref(M), ctypes.byref(opts), 12, px, px, None, *([null] * 12),
                                           None) == lib.KGEO_E_BADARG
    h = lib.PeerHandle()
    p = ctypes.c_void_p()
    assert L.kgeo_peer_alloc(0, ctypes.byref(p), ctypes.byref(h)) == lib.KGEO_E_BADARG
    assert L.kgeo_peer_open(None, ctypes.byref(p)) == lib.KGEO_E_BADARG
    assert L.kgeo_peer_close(None) == 0 and L.kgeo_peer_free(None) == 0


def test_no_cpu_fallback():
    import torch
    import kgeo_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kgeo_b200.raytrace_ana(0.94, [0, 1000., 0.3, 0], [np.zeros(3), np.ones(3)], plotdata=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kgeo_b200.make_image(0.94, np.inf, 0.3, 2, -5., 5., -5., 5., 1.)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kgeo_b200.special.elliprf(1.0, 2.0, 3.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kgeo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and False, "%s mentions the oracle" % f
                assert "scipy.special" not in text or "import scipy" not in text, "%s imports scipy" % f


def test_public_checks_match_reference_messages():
    import kgeo_b200 as kg
    A = np.zeros(2)
    for fn in (lambda: kg.raytrace_ana(1, [0, 1000., 0.3, 0], [A, A], plotdata=False),
               lambda: kg.coords_at_tau(2.0, [0, 1000., 0.3, 0], [A, A], np.zeros((1, 2))),
               lambda: kg.make_image(1, np.inf, 0.3, 2, -5., 5., -5., 5., 1.),
               lambda: kg.r_equatorial(-1.5, np.inf, 0.3, 0, A, A)):
        with pytest.raises(Exception, match=r"\|a\| should be a float in range \[0,1\)"):
            fn()
    with pytest.raises(Exception, match="r_o should be a float > 100"):
        kg.Iobs(0.5, 50., 0.3, 0, A, A)
    with pytest.raises(Exception, match=r"th_o should be a float in range \(0,pi/2\) or \(pi/2,pi\)"):
        kg.nmax_equatorial(0.5, 1000., 0.0, A, A)
    with pytest.raises(Exception, match="mbar should be an integer >=0!"):
        kg.Iobs(0.5, 1000., 0.3, 1.0, A, A)
    with pytest.raises(Exception, match="alpha, beta are different lengths!"):
        kg.r_equatorial(0.5, 1000., 0.3, 0, A, np.zeros(3))


def test_pixel_grid_quirk_and_defaults():
    import kgeo_b200 as kg
    from kgeo_b200 import equatorial_images as ei, kerr_raytracing_ana as ana
    A, B = ei.pixel_grid(-10., 10., -10., 10., 20. / 256)
    assert A.size == 256 * 256
    assert A[0] == -10. and np.isclose(A[255], 10.) and B[0] == B[255] == -10.     # row = constant beta
    A, B = ei.pixel_grid(-10.5, 10., -10., 10., 0.5)          # int(floor(20.5)/0.5) = 40 columns, as the reference
    assert A.size == 40 * 40
    assert (ana.SPIN, ana.ROUT, ana.NGEO, ana.NPIX) == (0.94, 1.e5, 250, 1000) and ana.MAXTAUFRAC == 1. - 1.e-10
    assert ei.SPECIND == 1 and ei.DISKANGLE == 0.1 and ei.SINTHB == np.pi / 4 and ei.OBSFREQ == 230.e9
    assert kg.MINSPIN == 1e-4 and kg.EP == 1e-12
    assert np.array_equal(kg.my_sign(np.array([-2., 0., 3.])), [-1., 1., 1.])


def test_setup_mode_flags():
    from kgeo_b200 import _engine, _lib
    th = 20 * np.pi / 180
    al = np.linspace(-6, 6, 1000)
    mode, clamped = _engine.setup_mode(0.94, th, al, 0 * al, True)     # reference default: vortical rays for |alpha|<a
    assert mode == (_lib.MODE_ANY_ETA_POS | _lib.MODE_ANY_ETA_NONPOS) and not clamped
    mode, clamped = _engine.setup_mode(0.94, th, np.array([3., 4.]), np.array([1., 1.]), True)
    assert mode == _lib.MODE_ANY_ETA_POS
    mode, clamped = _engine.setup_mode(0.0, th, np.array([0.0, 1.0]), np.array([0.0, 1.0]), True)   # eta == 0 pixel, spin 0
    assert mode == (_lib.MODE_ANY_ETA_POS | _lib.MODE_ANY_ETA_NONPOS) and not clamped
    mode, clamped = _engine.setup_mode(5e-5, th, np.array([0.0, 1.0]), np.array([0.0, 1.0]), True)  # eta<0 clamped to EP
    assert clamped and mode == (_lib.MODE_ANY_ETA_POS | _lib.MODE_CLAMP_LOWSPIN_ETA)


def test_model_packing():
    import kgeo_b200 as kg
    from kgeo_b200 import _lib
    from kgeo_b200.models import pack_models
    M = pack_models(kg.Emissivity('ring', emiscut_in=3.), kg.Velocity('general', fac_subkep=0.7, beta_phi=0.6, beta_r=0.5),
                    kg.Bfield('bz_guess', C=2.))
    assert (M.emissivity_type, M.velocity_type, M.bfield_type) == (_lib.EM['ring'], _lib.VEL['general'], _lib.BF['bz_guess'])
    assert list(M.velocity_p)[:4] == [0.0, 0.7, 0.6, 0.5]
    assert list(M.emissivity_p)[:3] == [4.5, 0.0, 0.3] and M.emiscut_in == 3. and M.emiscut_out == 1.e10
    assert M.bfield_p[0] == 2.
    M = pack_models(kg.Emissivity('glm'), kg.Velocity('kep', retrograde=True), kg.Bfield('vert'))
    assert list(M.emissivity_p)[:2] == [-1.5, 0.5] and M.velocity_p[0] == 1.0 and list(M.bfield_p)[:3] == [0., 1., 0.]
    M = pack_models(kg.Emissivity('thermal', ne0=1e5, Te0=1e11, B0=5., bfield_model=True))
    assert list(M.emissivity_p) == [5.0, 1e5, 1e11, 5.0, 0.7, 1.0, 1.5, 1.0]
    b = kg.Bfield('power', p=0.75, C=2., usemono=True)
    assert b.descriptor() == (_lib.BF['power'], [0.75, 2.0, 1.0])
    assert kg.Bfield('bz_monopole', secondorder_only=True).descriptor() == (_lib.BF['bz_monopole'], [1.0, 1.0])
    v = kg.Velocity('driftframe', bfield=b, nu_parallel=0.25)
    assert v.descriptor() == (_lib.VEL['driftframe'], [0.25]) and kg.Velocity('driftframe').bfield.fieldtype == 'bz_para'
    M = pack_models(velocity=v)                       # the drift frame brings its own field model along
    assert M.bfield_type == _lib.BF['power'] and list(M.bfield_p)[:3] == [0.75, 2.0, 1.0]
    with pytest.raises(NotImplementedError):
        pack_models(velocity=v, bfield=kg.Bfield('bz_para'))
    with pytest.raises(NotImplementedError):
        kg.Bfield('const_comoving')
    with pytest.raises(NotImplementedError):
        kg.Velocity('driftframe', nu_parallel='FF')
    with pytest.raises(Exception, match="nu_parallel should be in the range"):
        kg.Velocity('driftframe', nu_parallel=1.5)


def test_row_ranges():
    from kgeo_b200 import sharding
    for nrows in (1, 7, 256, 4096, 1000):
        for world in (1, 2, 3, 4, 8):
            rr = sharding.all_row_ranges(nrows, world)
            assert rr[0][0] == 0 and rr[-1][1] == nrows
            assert all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
            sizes = [b - a for a, b in rr]
            assert max(sizes) - min(sizes) <= 1
    A, B = np.meshgrid(np.arange(5.), np.arange(7.))
    al, be, (r0, r1) = sharding.shard_pixels(A.flatten(), B.flatten(), 5, 1, 3)
    assert (r0, r1) == (3, 5) and np.array_equal(np.unique(be), [3., 4.]) and al.size == 10
    with pytest.raises(ValueError):
        sharding.row_range(10, 3, 3)


def test_solver_host_side():
    """kgeo_b200.solver: stream-function closed forms (host helpers, as in kgeo/solver.py:13-37) and the model ids of the ABI"""
    from kgeo_b200 import solver
    from oracle import kgeo_oracle as ko
    r, th = np.linspace(1.5, 40., 50), np.linspace(0.05, 3.1, 50)
    for model, pval in (('para', 1), ('mono', 1), ('power', 0.75)):
        assert np.allclose(solver.psifunc(r, th, 0.7, model=model, pval=pval), ko.psifunc(r, th, 0.7, model, pval), rtol=1e-15)
    assert solver.psifunc(r, th, 0.7, model='nonsense') == 0
    assert [solver._model_id(m) for m in ('para', 'mono', 'power', 'nonsense')] == [0, 1, 2, 3]
    assert solver.MAXTAUFRAC == 1. - 1.e-10


def test_jet_host_scalars(golden):
    """kgeo_b200.off_eq: the one-per-call host scalars of getstokes -- field-line speed at the first crossing, stagnation radius
    (quintic / nested secant solves restating SciPy's scalar `newton`), power-law normalisation by Gauss-Legendre -- against the
    checker's SciPy versions on the crossings the unmodified reference found (tests/golden/jet_pipeline.npz)."""
    import scipy.integrate as scint
    from kgeo_b200 import off_eq
    from oracle import kgeo_oracle as ko
    g = golden('jet_pipeline')
    for name in ('mono', 'power'):
        a, r_o, th_o, psit, pval, neqmax, ngeo = [float(x) for x in g[name + '_par']]
        shape = tuple(g[name + '_shape'])
        r, th = g[name + '_root_r'], g[name + '_root_th']
        field, r0, norm = off_eq.jet_scalars(name, r, th, shape, psit, a, 2, pval)
        rn, thn = r.reshape(shape), np.arccos(np.abs(np.cos(th.reshape(shape))))
        i0 = np.where(np.nan_to_num(rn)[0] != 0)[0][0]
        with np.errstate(all='ignore'):
            if name == 'mono':
                om = ko.bfield_lab_omega(dict(type='bz_monopole'), a, rn[0][i0:i0 + 1], thn[0][i0])[3][0]
                assert field == off_eq.JET_MONO and norm == 1.0 and abs(r0 - ko.r0min_mono(thn[0][i0], om, a)) <= 1e-13 * r0
            else:
                om = np.broadcast_to(ko.bfield_lab_omega(dict(type='power', p=pval), a, rn[0][i0:i0 + 1], thn[0][i0])[3], (1,))[0]
                r0o, _ = ko.r0min_power(psit, om, a, pval)
                want = scint.quad(lambda x: np.exp(-(x - r0o) ** 2 / 8) * (x ** 2 + a ** 2 * (1 - psit / x ** pval) ** 2), r0o, np.inf,
                                  epsabs=1e-13, epsrel=1e-13)[0]
                assert field == off_eq.JET_POWER and abs(r0 - r0o) <= 1e-13 * r0o and abs(norm - want) <= 1e-12 * want
    with pytest.raises(NotImplementedError):
        off_eq.getstokes(1.0, np.zeros((2, 2)), np.zeros((2, 2)), 1000., 0.3, 0.5, 50, model='para')


def test_savegeos_loadgeos_roundtrip(tmp_path):
    """Geodesics.savegeos / loadgeos with the reference's HDF5 dataset names (kgeo/kerr_raytracing_utils.py:201-240), on a
    host-constructed Geodesics object (no GPU).  Skipped when h5py is not installed; without it both raise ImportError."""
    import numpy as np
    from kgeo_b200 import kerr_raytracing_utils as ku
    rng = np.random.default_rng(3)
    arrs = [rng.normal(size=(5, 7)) for _ in range(6)]
    g = ku.Geodesics(0.5, [0, 1000., 0.3, 0], [rng.normal(size=7), rng.normal(size=7)], arrs[0], arrs[1], arrs[2:])
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            g.savegeos(str(tmp_path / "x.h5"))
        with pytest.raises(ImportError):
            ku.loadgeos(str(tmp_path / "x.h5"))
        pytest.skip("h5py is not installed")
    fname = g.savegeos(str(tmp_path / "x.h5"))
    with h5py.File(fname, 'r') as hf:
        assert sorted(hf.keys()) == sorted(['spin', 'inc', 'r_o', 'alpha', 'beta', 't', 'r', 'theta', 'phi', 'affine', 'mino'])
    g2 = ku.loadgeos(fname)
    assert g2.a == g.a and g2.r_o == g.r_o and g2.th_o == g.th_o
    for x, y in zip(g.geo_coords + [g.mino_times, g.affine_times], g2.geo_coords + [g2.mino_times, g2.affine_times]):
        assert np.array_equal(x, y)


def test_ncu_summary_belongs_to_the_current_kernels():
    """bench.py takes its roofline constants (executed flops per point / pixel, pipe activity, DRAM traffic) from
    profiles/r02_ncu_summary.json only while the CUDA sources hash to the value recorded with the capture.  A mismatch is not an
    error (the bench then prints null) but it means the capture has to be redone: reported as a skip."""
    import json
    import bench
    summ = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_summary.json")))
    if summ.get("csrc_sha256_16") != bench.kernel_source_hash():
        pytest.skip("kgeo_b200/csrc changed since the ncu capture (%s): re-run tools/gpu_session23.sh" % summ.get("git_sha"))
    assert bench.ncu_summary() is not None
    assert 400 < summ["track_pass"]["flop_per_point"] < 800 and summ["track_pass"]["registers_hot"] == 64
