"""GPU parity: equatorial lensing, crossing counts, physics closed forms and the fused image kernel through the
public drop-in API vs fixtures of the unmodified reference, the oracle, and the reference's own ipole tests."""
import json

import numpy as np
import pytest

from helpers import assert_close, image_models, IMAGE_RTOL, TRACK_RTOL

pytestmark = pytest.mark.gpu
np.seterr(all='ignore')

IMAGES = ['default', 'testpol', 'subkep_pol', 'kep_vert_path', 'thermal_genpower', 'infall_tor', 'gelles_simple',
          'simfit_rm1', 'spin0_default']


@pytest.fixture(scope="module")
def kg():
    import kgeo_b200
    return kgeo_b200


def _models(kg, spec_json):
    em, vel, bf, kw = image_models(spec_json)
    pop = lambda d: {k: v for k, v in d.items() if k != 'type'}
    return (kg.Emissivity(em['type'], **pop(em)), kg.Velocity(vel['type'], **pop(vel)), kg.Bfield(bf['type'], **pop(bf)), kw)


def test_lensing_against_reference(golden, kg):
    g = golden('lensing')
    for i in range(int(g['ncases'])):
        a, th_o, r_o = [float(x) for x in g['c%d_par' % i]]
        A, B = g['c%d_alpha' % i], g['c%d_beta' % i]
        allr, allIr, allN, Imax_all = kg.r_equatorial_orders(a, r_o, th_o, 2, A, B)
        for mbar in range(3):
            rs, Ir, Imax, Nmax = kg.r_equatorial(a, r_o, th_o, mbar, A, B)
            ref = g['c%d_m%d' % (i, mbar)]
            assert np.array_equal(Nmax, ref[3]), "Nmax differs (case %d, m=%d)" % (i, mbar)
            assert_close(rs, ref[0], TRACK_RTOL, floor=1e-6, what='r_s %d/%d' % (i, mbar))
            assert_close(Ir, ref[1], 1e-13, what='Ir %d/%d' % (i, mbar))
            assert_close(Imax, ref[2], 2e-10, what='Imax %d/%d' % (i, mbar))
            # the fused all-orders pass gives the same bits as the per-order calls
            assert np.array_equal(allr[mbar], rs) and np.array_equal(allIr[mbar], Ir) and np.array_equal(allN[mbar], Nmax)
        key = 'c%d_cross0' % i
        if key in g.files:   # t, r, theta, phi at the first crossing (scripts/calc_t_equatorial.py)
            Ir = g['c%d_m0' % i][1]
            sig, (t, r, th, ph) = kg.coords_at_tau(a, [0, r_o, th_o, 0], [A, B], Ir)
            got = np.array([sig[0], t[0], r[0], th[0], ph[0]])
            assert_close(got, g[key], TRACK_RTOL, floor=1.0, what='crossing coords %d' % i)


def test_counts_against_reference(golden, kg):
    g = golden('counts')
    a, th_o = [float(x) for x in g['par']]
    neq = kg.n_equatorial_crossings(a, th_o, g['alpha'], g['beta'], g['tau'])
    npol = kg.n_poloidal_orbits(a, th_o, g['alpha'], g['beta'], g['tau'])
    assert np.array_equal(neq, g['neq'])
    assert_close(npol, g['npol'], 1e-13, what='npol')
    assert np.array_equal(kg.n_equatorial_crossings(a, th_o, g['alpha'], g['beta'], g['tau'][2]), g['neq'][2])


def test_angular_turnings_against_reference(golden, kg):
    """n_angular_turnings (kerr_raytracing_utils.py:686-792), ordinary and vortical rays, 2-D and 1-D tau, an equatorial observer."""
    g = golden('turnings')
    for i in range(int(g['ncases'])):
        a, th_o = [float(x) for x in g['c%d_par' % i]]
        A, B, tau = g['c%d_alpha' % i], g['c%d_beta' % i], g['c%d_tau' % i]
        got = kg.n_angular_turnings(a, th_o, A, B, tau)
        assert got.shape == tau.shape
        assert_close(got, g['c%d_n' % i], 1e-10, floor=1e-3, what='turnings %d' % i)
        got1 = kg.n_angular_turnings(a, th_o, A, B, tau[3])
        assert got1.shape == (len(A),)
        assert_close(got1, g['c%d_n1' % i], 1e-10, floor=1e-3, what='turnings %d (1-D tau)' % i)
    with pytest.raises(Exception, match=r"range \(0,pi/2\]"):
        kg.n_angular_turnings(0.5, 2.0, np.zeros(2), np.zeros(2), np.zeros((1, 2)))


def test_module_level_model_functions_against_reference(golden, kg):
    """u_zamo ... u_general (velocities.py:179-470) and emisBPL / emisGLM (emissivities.py:121-139) called as module functions."""
    from kgeo_b200 import velocities as kv, emissivities as ke
    g = golden('model_functions')
    for a in (0.0, 0.5, 0.94):
        tag = 'a%02d' % int(round(100 * a))
        r = g[tag + '_r']
        calls = dict(zamo=kv.u_zamo(a, r), infall=kv.u_infall(a, r), kep=kv.u_kep(a, r), kep_retro=kv.u_kep(a, r, retrograde=True),
                     subkep=kv.u_subkep(a, r, fac_subkep=0.7), gelles=kv.u_gelles(a, r, beta=0.4, chi=-2.0), simfit=kv.u_grmhd_fit(a, r),
                     general=kv.u_general(a, r, fac_subkep=0.7, beta_phi=0.6, beta_r=0.8))
        for k, u in calls.items():
            assert_close(np.array(u), g[tag + '_u_' + k], 1e-10, floor=1e-12, what='%s %s' % (tag, k))
        for k, j in (('bpl', ke.emisBPL(a, r)), ('bpl2', ke.emisBPL(a, r, p1=-1.5, p2=-0.3)), ('glm', ke.emisGLM(a, r)),
                     ('ring', ke.emisGLM(a, r, gamma_off=0.2, sigma=0.7, mu_ring=4.5))):
            assert_close(j, g[tag + '_' + k], 1e-10, floor=1e-300, what='%s %s' % (tag, k))
        from kgeo_b200 import bfields as kb
        th = np.pi / 2
        bcalls = dict(simple=kb.Bfield_simple(a, r, (0.87, 0.3, -0.5)), simple_rm1=kb.Bfield_simple_rm1(a, r, (0.4, -0.7, 0.2)))
        if a > 0:
            bcalls.update(bzmagic=kb.Bfield_BZmagic(a, r, th, C=-1.5), gen_power=kb.Bfield_gen_power(a, r, th, 0.5, 0.8, True))
        for k, b in bcalls.items():
            ref = g[tag + '_B_' + k]
            assert len(b) == ref.shape[0]
            assert_close(np.array(b), ref, 1e-10, floor=1e-6 * np.max(np.abs(ref)), what='%s B %s' % (tag, k))
    with pytest.raises(Exception, match="should be a float in range"):
        kv.u_kep(2, r)


def test_models_against_reference(golden, kg):
    g = golden('models')
    spec = json.loads(str(g['spec']))
    genvel = kg.Velocity('general', fac_subkep=0.6, beta_phi=0.8, beta_r=0.7)
    for ia, a in enumerate(g['spins']):
        a = float(a)
        r = g['a%d_r' % ia]
        for iv, (vt, kw) in enumerate(spec['vels']):
            got = np.array(kg.Velocity(vt, **kw).u_lab(a, r))
            assert_close(got, g['a%d_v%d' % (ia, iv)], 1e-12, what='vel %s a=%g' % (vt, a))
        for ib, (bt, kw) in enumerate(spec['bfs']):
            bf = kg.Bfield(bt, **kw)
            assert_close(np.array(bf.bfield_lab(a, r)), g['a%d_b%d' % (ia, ib)], 1e-12, what='bfield %s a=%g' % (bt, a))
            assert_close(bf.bsq(a, r, genvel), g['a%d_bsq%d' % (ia, ib)], 1e-12, what='bsq %s a=%g' % (bt, a))
        for ie, (et, kw) in enumerate(spec['ems']):
            got = kg.Emissivity(et, **kw).jrest(a, r, g=0.7 + 0 * r, sinthetab=0.6 + 0 * r, specind=1.0)
            assert_close(got, g['a%d_e%d' % (ia, ie)], 1e-12, what='emis %s a=%g' % (et, a))
        got = kg.Emissivity('thermal', ne0=1.e5, Te0=1.e11, B0=5.).jrest(a, r, g=0.7 + 0 * r, sinthetab=0.6 + 0 * r,
                                                                      Bmag=3.0 + 0 * r)
        assert_close(got, g['a%d_ethermal' % ia], 1e-12, what='thermal a=%g' % a)


def test_ipole_model_known_answers(kg):
    """tests/test_velocities.py:77-104, tests/test_bfields.py:8-26, tests/test_emissivities.py:5-43 of the reference."""
    vt = {(0, 1, 0.7, 0.7, 0.3, 4.0): [1.65511, -0.243725, 0, 0.141355],
          (0, 0.3, 0.1, 0.4, 0.9375, 1.6): [14.0203, -0.934042, 0, 3.63511],
          (0, 0.1, 0.36, 0.72, 0.1707, 3.4): [2.04774, -0.549666, 0, 0.0269644],
          (1, 0.24, 0.42, 0.24, 0.42, 10.): [1.19678, -0.340182, 0, -0.00297456],
          (1, 0.24, 0.42, 0.24, 0.42, 3.5): [2.23166, -0.71236, 0, 0.0121293]}
    for (retro, fs, bp, br, a, r), want in vt.items():
        v = kg.Velocity('general', retrograde=bool(retro), fac_subkep=fs, beta_phi=bp, beta_r=br)
        assert np.allclose(np.array(v.u_lab(a, r)).ravel(), want, rtol=1e-5)
    bt = {(1, 0.5, 2.9): -0.201073, (1, 0.5, 3.7): -0.141622, (1, 0.5, 10.): -0.0832697, (1, 0.9375, 1.35): -382.704,
          (1, 0.9375, 2.8): -0.546836, (-1, 0.9375, 2.8): -0.546836}
    for (C, a, r), want in bt.items():
        B = kg.Bfield('bz_guess', C=C).bfield_lab(a, r)
        assert np.allclose(want, B[2] / B[0], rtol=1e-5)
    glm = dict(gamma_off=-2, sigma=1.)
    ring = dict(r_ring=4, gamma_off=0.1, sigma=0.6)
    et = [('constant', 0, 1, {}, 1), ('bpl', 0.3, 8, {}, 0.022089), ('bpl', 0.3, 25, {}, 0.00023715),
          ('bpl', 0.4, 3, {}, 0.36913), ('bpl', 0.5, 10, {}, 0.0085085), ('bpl', 0.9, 3.141593, {}, 0.15376),
          ('bpl', 0.3679, 3.141593, {}, 0.33512), ('glm', 0.3679, 3.141593, {}, 0.19198), ('glm', 0.5, 5, {}, 0.069270),
          ('glm', 0.3, 5, {}, 0.066282), ('glm', 0.83, 9, {}, 0.014741), ('glm', 0.3679, 3.1415927, glm, 0.30569),
          ('glm', 0.4, 5, glm, 0.19078), ('ring', 0.3679, 3.141593, {}, 0.061782), ('ring', 0.4, 5, {}, 0.75227),
          ('ring', 0.83, 9, {}, 0.00067951), ('ring', 0.5, 2.4, ring, 0.16079), ('ring', 0.318, 7, ring, 0.017807),
          ('ring', 0.1, 5., ring, 0.32917)]
    for (model, a, r, kw, want) in et:
        assert np.isclose(kg.Emissivity(model, **kw).jrest(float(a), np.array([float(r)])), want, rtol=1e-4)


def test_velocity_normalisation(kg):
    """u.u = -1 for every closed-form model (tests/test_velocities.py:26-53 of the reference, seed 42)."""
    rng = np.random.default_rng(42)
    for vt, kw in (('zamo', {}), ('infall', {}), ('kep', {}), ('subkep', dict(fac_subkep=0.7)),
                   ('general', dict(fac_subkep=0.6, beta_phi=0.8, beta_r=0.7)), ('gelles', {})):
        for _ in range(5):
            a = float(rng.uniform(0.01, 0.99))
            rh = 1 + np.sqrt(1 - a * a)
            r = rng.uniform(rh * 1.05, 30, 64)
            u0, u1, u2, u3 = kg.Velocity(vt, **kw).u_lab(a, r)
            D = r * r - 2 * r + a * a
            g00, g11, g33, g03 = -(1 - 2 / r), r * r / D, r * r + a * a + 2 * a * a / r, -2 * a / r
            usq = g00 * u0 * u0 + g11 * u1 * u1 + g33 * u3 * u3 + 2 * g03 * u0 * u3 + r * r * u2 * u2
            assert np.allclose(usq, -1, rtol=1e-5), vt


@pytest.mark.parametrize('name', IMAGES)
def test_images_against_reference(golden, kg, name):
    g = golden('image_' + name)
    a, r_o, th_o, mbar_max, half, npx = [float(x) for x in g['par']]
    em, vel, bf, kw = _models(kg, g['spec'])
    out = kg.make_image(a, r_o, th_o, int(mbar_max), -half, half, -half, half, 2 * half / npx, emissivity=em, velocity=vel,
                        bfield=bf, **kw)
    names = ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np', 'lp', 'Ie'][:len(out)]
    assert len(out) == (11 if kw.get('pathlength') else 9)
    ref_I = g['I']
    for got, key in zip(out, names):
        ref = g[key]
        assert got.shape == ref.shape, key
        assert np.array_equal(got == 0, ref == 0), "%s %s: zero mask differs" % (name, key)
        if key in ('Q', 'U'):      # Q, U change sign: gate relative to the pixel's total intensity
            den = np.where(ref_I != 0, np.abs(ref_I), 1.0)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), key
            assert np.nanmax(np.abs(got - ref) / den) <= IMAGE_RTOL, key
        elif key in ('n',):
            assert np.array_equal(got, ref)
        else:
            assert_close(got, ref, IMAGE_RTOL if key in ('I', 'g', 'sinthb', 'lp', 'Ie', 'np') else TRACK_RTOL,
                         floor=1e-300, what='%s %s' % (name, key))


def test_Iobs_single_order(golden, kg):
    g = golden('image_kep_vert_path')
    a, r_o, th_o, mbar_max, half, npx = [float(x) for x in g['par']]
    em, vel, bf, kw = _models(kg, g['spec'])
    A, B = kg.equatorial_images.pixel_grid(-half, half, -half, half, 2 * half / npx)
    res = kg.Iobs(a, r_o, th_o, 1, A, B, emissivity=em, velocity=vel, bfield=bf, **kw)
    assert len(res) == 11
    I, Q, U, gg, rs, sinthb, Ir, Imax, Nmax, lp, Ie = res
    assert_close(I, g['I'][:, 1], IMAGE_RTOL, floor=1e-300, what='Iobs I')
    assert_close(rs, g['r'][:, 1], TRACK_RTOL, floor=1e-300, what='Iobs r')
    assert_close(lp, g['lp'][:, 1], IMAGE_RTOL, floor=1e-300, what='Iobs lp')
    r2, Ir2, Imax2, Nmax2 = kg.r_equatorial(a, r_o, th_o, 1, A, B)
    assert np.array_equal(Imax, Imax2) and np.array_equal(Ir, Ir2)


def _evpa(Q, U):
    evpa = 180. / np.pi * 0.5 * np.arctan2(U, Q) + 90.
    evpa[evpa > 90] -= 180.
    evpa[evpa > 90] -= 180.
    return evpa


IPOLE = [('lowspin_keplerian', (0.01, 0.1, 1.0, 1.0, 1.0), [-10, -3, 3, 10], [-6.05, -5.25, 5.25, 6], (1e-2, 4e-2, 1e-2, 1e-2)),
         ('midspin_inclined_keplerian', (0.42, 60., 1.0, 1.0, 1.0), [-10, -4.3125, 5., 10.], [-5.3, -4.4, 5.7, 6.8], (1e-2, 3e-2, 1e-2, 1e-2)),
         ('midspin_inclined_subkep', (0.7, 45., 0.6, 0.7, 0.8), [-10, -2.7, 4, 10], [-5., -4, 5.8, 7.], (1e-2, 2e-2, 1e-2, 1e-2))]


@pytest.mark.parametrize('name,par,n0_limits,n1_limits,rtols', IPOLE)
def test_ipole_evpa(kg, name, par, n0_limits, n1_limits, rtols):
    """The reference's tests/test_images.py:62-123: EVPA cuts of 800x800 polarized n=0,1 images vs ipole."""
    import os
    from conftest import GOLDEN
    bhspin, inc, fac_subkep, beta_r, beta_phi = par
    nx = 800
    ploc, n0_a, n0_b, n1_a, n1_b = np.loadtxt(os.path.join(GOLDEN, 'ipole', name + '.csv'), delimiter=',', unpack=True)
    out = kg.make_image(bhspin, np.inf, inc * np.pi / 180., 2, -10., 10., -10., 10., 20. / nx,
                        emissivity=kg.Emissivity('ring'),
                        velocity=kg.Velocity('general', fac_subkep=fac_subkep, beta_phi=beta_phi, beta_r=beta_r),
                        bfield=kg.Bfield('bz_guess', C=1.), polarization=True, specind=1.)
    Q, U = out[1], out[2]
    e0 = _evpa(Q[:, 0].reshape(nx, nx), U[:, 0].reshape(nx, nx))
    e1 = _evpa(Q[:, 1].reshape(nx, nx), U[:, 1].reshape(nx, nx))
    m0a = (ploc > n0_limits[0]) & (ploc < n0_limits[1])
    m0b = (ploc > n0_limits[2]) & (ploc < n0_limits[3])
    m1a = (ploc > n1_limits[0]) & (ploc < n1_limits[1])
    m1b = (ploc > n1_limits[2]) & (ploc < n1_limits[3])
    assert np.allclose(n0_a[m0a], e0[400, m0a], rtol=rtols[0])
    assert np.allclose(n0_b[m0b], e0[m0b, 400], rtol=rtols[1])
    assert np.allclose(n1_a[m1a], e1[400, m1a], rtol=rtols[2])
    assert np.allclose(n1_b[m1b], e1[m1b, 400], rtol=rtols[3])


def test_image_against_oracle_cfg2_and_shards(kg):
    """BASELINE cfg 2 (256x256, a=0.94, 17 deg, n=0..2, default models) vs the oracle; and row shards == whole image."""
    from oracle import kgeo_oracle as ko
    from kgeo_b200 import sharding, equatorial_images as ei
    a, r_o, th_o = 0.94, np.inf, 17 * np.pi / 180
    out = kg.make_image(a, r_o, th_o, 2, -10., 10., -10., 10., 20. / 256)
    ref = ko.make_image(a, r_o, th_o, 2, -10., 10., -10., 10., 20. / 256, dict(type='bpl'), dict(type='zamo'), dict(type='rad'))
    for got, want, key in zip(out, ref, ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np']):
        assert np.array_equal(got == 0, want == 0), key
        assert_close(got, want, IMAGE_RTOL, floor=1e-300, what='cfg2 ' + key)
    A, B = ei.pixel_grid(-10., 10., -10., 10., 20. / 256)
    for world in (2, 8):
        for rank in range(world):
            al, be, (r0, r1) = sharding.shard_pixels(A, B, 256, rank, world)
            part = ei.image_device(a, r_o, th_o, 0, 2, al, be, ei.emis_default, ei.vel_default, ei.bfield_default, False,
                                   False, ei.SPECIND, ei.DISKANGLE, ei.OBSFREQ)
            assert np.array_equal(part["I"].cpu().numpy().T, out[0][r0 * 256:r1 * 256], equal_nan=True)
            assert np.array_equal(part["r"].cpu().numpy().T, out[3][r0 * 256:r1 * 256], equal_nan=True)


def test_image_host_entry_point(kg):
    import ctypes
    from kgeo_b200 import _lib, _engine, equatorial_images as ei
    from kgeo_b200.models import pack_models
    a, r_o, th_o = 0.94, np.inf, 17 * np.pi / 180
    A, B = ei.pixel_grid(-10., 10., -10., 10., 20. / 64)
    out = kg.make_image(a, r_o, th_o, 2, -10., 10., -10., 10., 20. / 64)
    n = A.size
    host = [np.empty((3, n)) for _ in range(7)] + [np.empty(n) for _ in range(3)]
    obs = _engine.observer_struct(a, r_o, th_o)
    M = pack_models(ei.emis_default, ei.vel_default, ei.bfield_default)
    opts = _lib.ImageOpts(0, 2, 0, 0, float(ei.SPECIND), ei.DISKANGLE, ei.OBSFREQ)
    _lib.check(_lib.lib().kgeo_equatorial_image_host(ctypes.byref(obs), ctypes.byref(M), ctypes.byref(opts), n,
                                                     _engine.hptr(A), _engine.hptr(B), *[_engine.hptr(h) for h in host],
                                                     None, None, 0))
    for got, ref in zip(host[:7], out[:7]):
        assert np.array_equal(got.T, ref, equal_nan=True)
    assert np.array_equal(host[7], out[7]) and np.array_equal(host[8], out[8])


def test_image_argument_checks(kg):
    with pytest.raises(Exception, match="mbar_max should be an integer"):
        kg.make_image(0.5, np.inf, 0.3, 1.0, -5., 5., -5., 5., 1.)
    with pytest.raises(Exception, match="mbar should be an integer"):
        kg.r_equatorial(0.5, np.inf, 0.3, -1, np.zeros(2), np.zeros(2))
    with pytest.raises(Exception, match="not recognized"):
        kg.Velocity('nonsense')
    with pytest.raises(Exception, match="all field coefficients are 0"):
        kg.Bfield('simple')


@pytest.mark.parametrize("interleave,pol", [(True, False), (True, True), (False, False)])
def test_sharded_image_kernel_stores_rows_in_place(kg, interleave, pol):
    """kgeo_equatorial_image_sharded: three ragged row shards (interleaved or blocks of rows) written by
    three calls into ONE destination buffer (a PeerBuffer: kgeo_peer_alloc memory, here on the same GPU) == the whole image."""
    from kgeo_b200 import sharding, equatorial_images as ei
    a, r_o, th_o, world = 0.94, np.inf, 17 * np.pi / 180, 3
    kw = dict(polarization=pol, pathlength=pol)
    if pol:
        kw.update(emissivity=kg.Emissivity('ring'), velocity=kg.Velocity('general', fac_subkep=0.7, beta_phi=0.7, beta_r=0.7),
                  bfield=kg.Bfield('bz_guess', C=1.), specind=1.)
    whole = kg.make_image(a, r_o, th_o, 2, -10., 10., -10., 10., 20. / 50, **kw)
    A, B = ei.pixel_grid(-10., 10., -10., 10., 20. / 50)
    ncols = 50
    nrows = A.size // ncols
    off, total = sharding.plane_layout(3, A.size, pathlength=pol)
    buf = sharding.PeerBuffer(total + 7)
    try:
        view = buf.tensor()
        view.fill_(-7.0)
        model = sharding._image_kwargs(kw)
        for rank in range(world):
            if interleave:
                al, be, rows = sharding.shard_pixels_interleaved(A, B, ncols, rank, world)
                step, offset = world, rank
            else:
                al, be, rows = sharding.shard_pixels(A, B, ncols, rank, world)
                step, offset = 1, rows[0]
            ei.image_into(a, r_o, th_o, 0, 2, al, be, *model, dst={k: buf.ptr + 8 * o for k, o in off.items()}, cols=ncols,
                          row_step=step, row_offset=offset, plane_stride=A.size)
        flat = sharding.to_host_pinned(view)
        assert np.all(flat[total:] == -7.0)                      # nothing stored past the planes
        got = sharding._image_tuple(flat[:total], off, 3, A.size, pol)
        assert len(got) == len(whole)
        for g_, w_ in zip(got, whole):
            assert g_.shape == w_.shape and np.array_equal(g_, w_, equal_nan=True)
        del view
    finally:
        buf.close()
    assert nrows % world != 0                                    # the shards were ragged


def test_shard_scatter_places_rows_with_the_copy_engines(kg):
    """kgeo_shard_scatter: local planes [nplanes, npix] -> their rows of larger destination planes (interleaved rows: one 3-D copy;
    planes that are not whole pitches apart: per-plane 2-D copies) == the NumPy placement."""
    import ctypes
    import torch
    from kgeo_b200 import _lib, _engine
    L = _lib.lib()
    rng = np.random.default_rng(3)
    ncols, nplanes = 12, 5
    for rows_local, step, offset, nrows in ((4, 3, 1, 12), (4, 1, 5, 11), (4, 3, 1, 11)):
        src = rng.standard_normal((nplanes, rows_local * ncols))
        d_src = _engine.to_device(src)
        d_dst = torch.full((nplanes, nrows * ncols), -7.0, dtype=torch.float64, device=d_src.device)
        smap = _lib.ShardMap(ncols, step, offset, nrows * ncols, 0)
        _lib.check(L.kgeo_shard_scatter(_engine.dptr(d_src), nplanes, rows_local * ncols, ctypes.byref(smap), _engine.dptr(d_dst),
                                        _engine.stream_ptr()))
        want = np.full((nplanes, nrows, ncols), -7.0)
        want[:, offset + step * np.arange(rows_local), :] = src.reshape(nplanes, rows_local, ncols)
        assert np.array_equal(d_dst.cpu().numpy().reshape(nplanes, nrows, ncols), want)


def test_sharded_image_argument_checks(kg):
    import ctypes
    from kgeo_b200 import _lib, _engine, equatorial_images as ei
    from kgeo_b200.models import pack_models
    L = _lib.lib()
    obs = _engine.observer_struct(0.5, np.inf, 0.3)
    M = pack_models(ei.emis_default, ei.vel_default, ei.bfield_default)
    opts = _lib.ImageOpts(0, 2, 0, 0, 1.0, 0.1, 230e9)
    x = _engine.to_device(np.zeros(12))
    null = ctypes.cast(None, _lib.c_double_p)
    for smap in (_lib.ShardMap(5, 1, 0, 12, 0),       # 12 pixels are not whole rows of 5
                 _lib.ShardMap(4, 0, 0, 12, 0),       # row_step 0
                 _lib.ShardMap(4, 2, 1, 12, 0)):      # destination plane (rows 1, 3, 5 -> 24 pixels) larger than plane_stride
        rc = L.kgeo_equatorial_image_sharded(ctypes.byref(obs), ctypes.byref(M), ctypes.byref(opts), 12, _engine.dptr(x),
                                             _engine.dptr(x), ctypes.byref(smap), *([null] * 12), None)
        assert rc == _lib.KGEO_E_BADARG, L.kgeo_last_error()


def test_multi_gpu_sharded_image_and_tracks():
    """N-GPU path under torchrun + NCCL (skipped on a single-GPU box): tests/dist_check.py."""
    import os
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 8)),
                          "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert "dist_check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_spin_inclination_sweep_against_oracle(kg):
    """BASELINE cfg 5 grid: a in linspace(0, 0.99, 8) x theta_o in linspace(5, 85, 8) degrees (64 configs, default models,
    n = 0..2), at 24 x 24 pixels so that the oracle finishes in seconds.  a = 0 exercises the |a| < MINSPIN branches,
    theta_o = 85 deg the class-II region.  Integer outputs and masks exact, r_s to 1e-10, I and g to 1e-8."""
    from oracle import kgeo_oracle as ko
    worst = dict(r=0.0, I=0.0, g=0.0)
    for a in np.linspace(0, 0.99, 8):
        for inc in np.linspace(5, 85, 8):
            a, th_o = float(a), float(inc * np.pi / 180)
            out = kg.make_image(a, np.inf, th_o, 2, -9., 9., -9., 9., 18. / 24)
            ref = ko.make_image(a, np.inf, th_o, 2, -9., 9., -9., 9., 18. / 24, dict(type='bpl'), dict(type='zamo'), dict(type='rad'))
            names = ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np']
            for got, want, key in zip(out, ref, names):
                assert np.array_equal(np.isnan(got), np.isnan(want)), (a, inc, key)
                assert np.array_equal(got == 0, want == 0), "zero mask differs: a=%g inc=%g %s" % (a, inc, key)
                if key == 'n':
                    assert np.array_equal(got, want), (a, inc)
                    continue
                ok = np.isfinite(want) & (want != 0)
                den = np.abs(want[ok])
                if key == 'r':      # tau -> r has condition number ~ r/(r4 - r3): crossings at r ~ 1e6 M exist in this sweep
                    den = den * np.maximum(1.0, den / 1000.0)
                e = np.max(np.abs(got[ok] - want[ok]) / den) if ok.any() else 0.0
                if key in worst:
                    worst[key] = max(worst[key], e)
                tol = TRACK_RTOL if key in ('r', 't') else IMAGE_RTOL
                assert e <= tol, "a=%g inc=%g %s: %.2e" % (a, inc, key, e)
    print("sweep worst rel errors", worst)


# ---- tabulated 'fromfile' models -----------------------------------------------------------------
TABLE = None


def _table_file():
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'a0p6_inflow.csv')


def test_fromfile_models_against_reference(golden, kg):
    """Velocity('fromfile') / Bfield('fromfile'): element-wise u^mu, B^i, b^2 (incl. radii beyond both table ends) and one
    polarized image vs the unmodified reference (tests/golden/fromfile_models.npz)."""
    g = golden('fromfile_models')
    a, r = float(g['a']), g['r']
    v = kg.Velocity('fromfile', file=_table_file())
    b = kg.Bfield('fromfile', file=_table_file())
    assert_close(np.array(v.u_lab(a, r)), g['u'], 1e-12, what='fromfile u_lab')
    assert_close(np.array(b.bfield_lab(a, r)), g['B'], 1e-12, what='fromfile bfield_lab')
    assert_close(b.bsq(a, r, v), g['bsq'], 1e-11, what='fromfile bsq')
    a_, r_o, th_o, mbar_max, half, npx = [float(x) for x in g['par']]
    out = kg.make_image(a_, r_o, th_o, int(mbar_max), -half, half, -half, half, 2 * half / npx,
                        emissivity=kg.Emissivity('ring', r_ring=3.5, sigma=0.6), velocity=v, bfield=b, polarization=True,
                        specind=1.)
    ref_I = g['img_I']
    for got, key in zip(out, ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np']):
        ref = g['img_' + key]
        assert got.shape == ref.shape and np.array_equal(got == 0, ref == 0), key
        if key in ('Q', 'U'):
            den = np.where(ref_I != 0, np.abs(ref_I), 1.0)
            assert np.nanmax(np.abs(got - ref) / den) <= IMAGE_RTOL, key
        elif key == 'n':
            assert np.array_equal(got, ref)
        else:
            assert_close(got, ref, IMAGE_RTOL if key in ('I', 'g', 'sinthb', 'np') else TRACK_RTOL, floor=1e-300,
                         what='fromfile image %s' % key)


def test_fromfile_known_answers_of_the_reference_tests(kg):
    """tests/test_velocities.py:56-74 and tests/test_bfields.py:29-44 of the reference (same data file, same numbers)."""
    v = kg.Velocity('fromfile', file=_table_file())
    for radius, want in {2.15708: [5.0056, -0.54416, 0.0, 0.75899], 2.5: [3.0016, -0.45323831, 0.0, 0.41926807],
                         3.58635: [1.7755, -0.22549, 0.0, 0.19684], 4.90566: [1.4859, -0.021582, 0.0, 0.12754]}.items():
        assert np.allclose(np.squeeze(v.u_lab(0.6, np.array(radius))), want, rtol=1.e-4)
    b = kg.Bfield('fromfile', file=_table_file())
    for radius, want in {2.15708: [0.84877, 0.0, -0.53419], 2.5: [0.63206, 0.0, -0.23666],
                         3.58635: [0.30706, 0.0, -0.066873], 4.90566: [0.16411, 0.0, -0.029673]}.items():
        assert np.allclose(np.squeeze(b.bfield_lab(0.6, radius)), want, rtol=1.e-4)


def test_image_sweep_matches_single_images(kg):
    """BASELINE cfg 5 entry point without a process group (one GPU): every configuration's image == make_image."""
    from kgeo_b200 import sharding
    cfgs = [(0.0, 5 * np.pi / 180), (0.42, 45 * np.pi / 180), (0.99, 85 * np.pi / 180)]
    res = sharding.make_image_sweep_sharded(cfgs, np.inf, 2, -10., 10., -10., 10., 20. / 48)
    assert len(res) == 3
    for (a, th), got in zip(cfgs, res):
        want = kg.make_image(a, np.inf, th, 2, -10., 10., -10., 10., 20. / 48)
        assert len(got) == len(want)
        for g_, w_ in zip(got, want):
            assert g_.shape == w_.shape and np.array_equal(g_, w_, equal_nan=True)
