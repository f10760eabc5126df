"""Pins the CPU oracle (oracle/kgeo_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py) and against the reference's own
known-answer tests.  CPU only."""
import json

import numpy as np
import pytest

from oracle import kgeo_oracle as ko
from helpers import assert_close, image_models

np.seterr(all='ignore')
TIGHT = 1e-13   # the oracle follows the reference's operation order with the same SciPy


def test_pi_against_reference(golden):
    g = golden('special')
    assert_close(ko.legendre_pi(np.full(18, 0.8), g['pi_list_phi'], np.full(18, -0.33)), g['pi_list'], TIGHT, what='Pi list')
    assert_close(ko.legendre_pi(g['pi_n'], g['pi_phi'], g['pi_m']), g['pi'], 0.0, what='Pi sweep')
    assert_close(ko.legendre_pi(g['pisp_n'], g['pisp_phi'], g['pisp_m']), g['pisp'], TIGHT, what='Pi special')


def test_pi_against_mpmath():
    """The reference's print-only self check (scipy_ellip_binding.py:241-256) turned into an assertion."""
    import mpmath
    phis = [.026, .1, np.pi / 2., 2.37, 4.1, 6., 7.7, 14.3, 215.]
    for phi in phis + [-p for p in phis]:
        want = float(mpmath.re(mpmath.ellippi(0.8, phi, -0.33)))
        got = float(ko.legendre_pi(np.array([0.8]), np.array([phi]), np.array([-0.33]))[0])
        assert abs(got / want - 1) < 2e-13


def test_setup_against_reference(golden):
    g = golden('setup')
    for i in range(int(g['ncases'])):
        a, th_o, r_o = g['c%d_par' % i]
        A, B = g['c%d_alpha' % i], g['c%d_beta' % i]
        lam, eta = ko.conserved(a, th_o, A, B)
        up, um, ratio, a2um = ko.angular_consts(a, th_o, lam, eta)
        for got, key in ((up, 'up'), (um, 'um'), (ratio, 'ratio'), (a2um, 'a2um')):
            assert_close(got, g['c%d_%s' % (i, key)], 0.0, what='case %d %s' % (i, key))
        _, _, thp, thm, thclass = ko.angular_turning(a, th_o, lam, eta)
        assert_close(thp, g['c%d_thp' % i], 0.0, what='case %d theta_plus' % i)
        assert_close(thm, g['c%d_thm' % i], 0.0, what='case %d theta_minus' % i)
        assert np.array_equal(thclass, g['c%d_thclass' % i])
        r1, r2, r3, r4, region = ko.quartic_roots(a, lam, eta)
        roots = g['c%d_roots' % i]
        for got, ref in zip((r1, r2, r3, r4), roots):
            assert np.array_equal(got, ref)
        assert np.array_equal(region, g['c%d_rclass' % i])
        assert_close(ko.mino_total(a, r_o, eta, r1, r2, r3, r4), g['c%d_tmax' % i], 0.0, what='tau_tot %d' % i)
        assert np.array_equal(ko.outside_crit(a, th_o, A, B), g['c%d_outside' % i])
        assert_close(ko.nmax_equatorial(a, r_o, th_o, A, B), g['c%d_nmax_eq' % i], 0.0, what='nmax_eq %d' % i)
        assert_close(ko.nmax_poloidal(a, r_o, th_o, A, B), g['c%d_nmax_pol' % i], TIGHT, what='nmax_pol %d' % i)


@pytest.mark.parametrize('name', ['default', 'grid17', 'classII', 'extremal', 'spin0', 'south', 'midspin'])
def test_tracks_against_reference(golden, name):
    g = golden('tracks_' + name)
    a, th_o, r_o, ngeo = g['par']
    tau, sig, (t, r, th, ph) = ko.raytrace(float(a), [0, float(r_o), float(th_o), 0], [g['alpha'], g['beta']], int(ngeo))
    assert_close(tau, g['tau'], 0.0, what='tau')
    for got, key in ((sig, 'sig'), (t, 't'), (r, 'r'), (th, 'th'), (ph, 'ph')):
        assert_close(got, g[key], TIGHT, floor=1e-3, what='%s %s' % (name, key))


def test_lensing_against_reference(golden):
    g = golden('lensing')
    for i in range(int(g['ncases'])):
        a, th_o, r_o = g['c%d_par' % i]
        A, B = g['c%d_alpha' % i], g['c%d_beta' % i]
        for mbar in range(3):
            got = np.array(ko.r_equatorial(float(a), float(r_o), float(th_o), mbar, A, B))
            assert_close(got, g['c%d_m%d' % (i, mbar)], TIGHT, floor=1e-6, what='r_equatorial %d/%d' % (i, mbar))
        key = 'c%d_cross0' % i
        if key in g.files:
            Ir = g['c%d_m0' % i][1]
            sig, co = ko.coords_at_tau(float(a), [0, float(r_o), float(th_o), 0], [A, B], Ir)
            got = np.array([sig[0], co[0][0], co[1][0], co[2][0], co[3][0]])
            assert_close(got, g[key], TIGHT, floor=1e-3, what='crossing coords %d' % i)


IMAGES = ['default', 'testpol', 'subkep_pol', 'kep_vert_path', 'thermal_genpower', 'infall_tor', 'gelles_simple',
          'simfit_rm1', 'spin0_default']


@pytest.mark.parametrize('name', IMAGES)
def test_images_against_reference(golden, name):
    g = golden('image_' + name)
    a, r_o, th_o, mbar_max, half, npx = g['par']
    em, vel, bf, kw = image_models(g['spec'])
    out = ko.make_image(float(a), float(r_o), float(th_o), int(mbar_max), -half, half, -half, half, 2 * half / npx,
                        em, vel, bf, **kw)
    names = ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np', 'lp', 'Ie'][:len(out)]
    for got, key in zip(out, names):
        assert_close(got, g[key], 1e-12, floor=1e-300, what='%s %s' % (name, key))


def test_models_against_reference(golden):
    g = golden('models')
    spec = json.loads(str(g['spec']))
    genvel = dict(type='general', fac_subkep=0.6, beta_phi=0.8, beta_r=0.7)
    for ia, a in enumerate(g['spins']):
        a = float(a)
        r = g['a%d_r' % ia]
        for iv, (vt, kw) in enumerate(spec['vels']):
            got = np.array([np.broadcast_to(c, r.shape) for c in ko.velocity_u_lab(dict(type=vt, **kw), a, r)])
            assert_close(got, g['a%d_v%d' % (ia, iv)], TIGHT, what='vel %s a=%g' % (vt, a))
        for ib, (bt, kw) in enumerate(spec['bfs']):
            got = np.array([np.broadcast_to(c, r.shape) for c in ko.bfield_lab(dict(type=bt, **kw), a, r)])
            assert_close(got, g['a%d_b%d' % (ia, ib)], TIGHT, what='bfield %s a=%g' % (bt, a))
            assert_close(ko.bfield_bsq(dict(type=bt, **kw), genvel, a, r), g['a%d_bsq%d' % (ia, ib)], TIGHT,
                         what='bsq %s a=%g' % (bt, a))
        for ie, (et, kw) in enumerate(spec['ems']):
            got = ko.emissivity_jrest(dict(type=et, **kw), a, r, g=0.7 + 0 * r, sinthb=0.6 + 0 * r, specind=1.0)
            assert_close(got, g['a%d_e%d' % (ia, ie)], TIGHT, what='emis %s a=%g' % (et, a))
        got = ko.emissivity_jrest(dict(type='thermal', ne0=1.e5, Te0=1.e11, B0=5.), a, r, g=0.7 + 0 * r,
                                  sinthb=0.6 + 0 * r, Bmag=3.0 + 0 * r)
        assert_close(got, g['a%d_ethermal' % ia], TIGHT, what='thermal a=%g' % a)


def test_counts_against_reference(golden):
    g = golden('counts')
    a, th_o = g['par']
    assert_close(ko.n_equatorial(float(a), float(th_o), g['alpha'], g['beta'], g['tau']), g['neq'], 0.0, what='neq')
    assert_close(ko.n_poloidal(float(a), float(th_o), g['alpha'], g['beta'], g['tau']), g['npol'], TIGHT, what='npol')


# ----- the reference's own known-answer tests, run against the oracle -----------------------------
def test_ipole_velocity_general():
    """tests/test_velocities.py:77-104 of the reference."""
    tests = {
        (0, 1, 0.7, 0.7, 0.3, 4.0): [1.65511, -0.243725, 0, 0.141355],
        (0, 0.3, 0.1, 0.4, 0.9375, 1.6): [14.0203, -0.934042, 0, 3.63511],
        (0, 0.1, 0.36, 0.72, 0.1707, 3.4): [2.04774, -0.549666, 0, 0.0269644],
        (1, 0.24, 0.42, 0.24, 0.42, 10.): [1.19678, -0.340182, 0, -0.00297456],
        (1, 0.24, 0.42, 0.24, 0.42, 3.5): [2.23166, -0.71236, 0, 0.0121293],
    }
    for (retro, fs, bp, br, a, r), want in tests.items():
        vel = dict(type='general', retrograde=bool(retro), fac_subkep=fs, beta_phi=bp, beta_r=br)
        got = np.array([np.squeeze(c) for c in ko.velocity_u_lab(vel, a, np.array([r]))], dtype=float)
        assert np.allclose(got, want, rtol=1e-5)


def test_ipole_bfield_bz_guess():
    """tests/test_bfields.py:8-26 of the reference."""
    tests = {(1, 0.5, 2.9): -0.201073, (1, 0.5, 3.7): -0.141622, (1, 0.5, 10.): -0.0832697,
             (1, 0.9375, 1.35): -382.704, (1, 0.9375, 2.8): -0.546836, (-1, 0.9375, 2.8): -0.546836}
    for (C, a, r), want in tests.items():
        B = ko.bfield_lab(dict(type='bz_guess', C=C), a, np.array([r]))
        assert np.allclose(want, B[2] / B[0], rtol=1e-5)


def test_ipole_emissivity():
    """tests/test_emissivities.py:5-43 of the reference."""
    glm = dict(gamma_off=-2, sigma=1.)
    ring = dict(r_ring=4, gamma_off=0.1, sigma=0.6)
    tests = [('constant', 0, 1, {}, 1), ('bpl', 0.3, 8, {}, 0.022089), ('bpl', 0.3, 25, {}, 0.00023715),
             ('bpl', 0.4, 3, {}, 0.36913), ('bpl', 0.5, 10, {}, 0.0085085), ('bpl', 0.9, 3.141593, {}, 0.15376),
             ('bpl', 0.3679, 3.141593, {}, 0.33512), ('glm', 0.3679, 3.141593, {}, 0.19198), ('glm', 0.5, 5, {}, 0.069270),
             ('glm', 0.3, 5, {}, 0.066282), ('glm', 0.83, 9, {}, 0.014741), ('glm', 0.3679, 3.1415927, glm, 0.30569),
             ('glm', 0.4, 5, glm, 0.19078), ('ring', 0.3679, 3.141593, {}, 0.061782), ('ring', 0.4, 5, {}, 0.75227),
             ('ring', 0.83, 9, {}, 0.00067951), ('ring', 0.5, 2.4, ring, 0.16079), ('ring', 0.318, 7, ring, 0.017807),
             ('ring', 0.1, 5., ring, 0.32917)]
    for (model, a, r, kw, want) in tests:
        assert np.isclose(ko.emissivity_jrest(dict(type=model, **kw), a, np.array([float(r)])), want, rtol=1e-4)


def test_ipole_midplane_rays():
    """tests/test_geodesics_kerr_ana.py:47-75 of the reference (5 of the 15 rays; the GPU test runs all)."""
    tests = {(0.42, 60., -9., 0.): (-0.103693, -8.09784), (0.42, 60., 6., 0.): (877.845, -13.6213),
             (0.9, 10., 1., 0.): (-0.518830, -0.632352), (0.73, 35., 0.0001, -1.): (1.02201, 1.13478),
             (0.73, 35., 0.0001, 4.): (-2.11221, 1.02322)}
    for (a, inc, al, be), want in tests.items():
        _, _, (t, r, th, ph) = ko.raytrace(a, [0, 1000., inc / 180. * np.pi, 0], [np.array([al]), np.array([be])], 10000)
        r, th, ph = r[:, 0], th[:, 0], ph[:, 0]
        x, y, z = r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)
        if z[0] > z[1]:
            x, y, z = x[::-1], y[::-1], z[::-1]
        assert np.allclose([np.interp(0, z, x), np.interp(0, z, y)], want, rtol=1e-3)


def test_inverse_lensing_oracle_matches_reference(golden):
    """rho_of_req / rho_of_req_single / critical_curve (equatorial_lensing.py:226-298): same SciPy brentq, same objective."""
    g = golden('inverse_lensing')
    vp = g['varphis'][::4]
    for ic, (a, th0, req) in enumerate(g['cases']):
        for mbar in (0, 1, 2):
            _, rho, al, be = ko.rho_of_req(float(a), float(th0), float(req), mbar, vp)
            assert np.max(np.abs(rho - g['rho_%d_%d' % (ic, mbar)][::4])) <= 1e-12
            assert np.max(np.abs(al - g['alpha_%d_%d' % (ic, mbar)][::4])) <= 1e-12
        al, be = ko.critical_curve(float(a), float(th0), 2001)
        assert np.array_equal(al, g['crit_alpha_%d' % ic]) and np.array_equal(be, g['crit_beta_%d' % ic])
    for m in (0, 1, 2):
        assert abs(ko.rho_of_req_single(0.94, 17 * np.pi / 180, 5.0, 0.3, mbar=m) - g['single'][m]) <= 1e-12


def test_fromfile_models_oracle_matches_reference(golden):
    """np.interp table models (kgeo/velocities.py:626-698, kgeo/bfields.py:749-806) incl. one polarized image."""
    import os
    g = golden('fromfile_models')
    F = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'a0p6_inflow.csv')
    a, r = float(g['a']), g['r']
    vel, bf = dict(type='fromfile', file=F), dict(type='fromfile', file=F)
    assert_close(np.array(ko.velocity_u_lab(vel, a, r)), g['u'], 1e-14, what='fromfile u')
    assert_close(np.array(ko.bfield_lab(bf, a, r)), g['B'], 1e-14, what='fromfile B')
    assert_close(ko.bfield_bsq(bf, vel, a, r), g['bsq'], 1e-13, what='fromfile bsq')
    a_, ro, th, mb, half, npx = g['par']
    img = ko.make_image(a_, ro, th, int(mb), -half, half, -half, half, 2 * half / npx, dict(type='ring', r_ring=3.5, sigma=0.6),
                        vel, bf, polarization=True, specind=1.)
    for k, val in zip(['I', 'Q', 'U', 'r', 't', 'g', 'sinthb'], img):
        assert_close(val, g['img_' + k], 1e-12, floor=1e-300, what='fromfile image ' + k)


# ---- BZ field models and the drift-frame velocity ---------------------------------------------------
def _bz_cases(g):
    import json
    spec = json.loads(str(g['spec']))
    for ia, a in enumerate(g['spins']):
        for ib, (bt, kw) in enumerate(spec['bfs']):
            yield ia, float(a), ib, dict(type=bt, **kw)


def test_bz_models_oracle_against_reference(golden):
    """Bfield('bz_monopole' | 'bz_para' | 'power' | ...) B^i, omega_field, b^2 and Velocity('driftframe') u^mu of the oracle
    vs the unmodified reference (tests/golden/bz_models.npz).  bz_para is complex-valued in the reference (scipy's
    lambertw): u^mu is compared where the reference's result is real (the fixture records that)."""
    g = golden('bz_models')
    vel = dict(type='general', fac_subkep=0.6, beta_phi=0.8, beta_r=0.7)
    with np.errstate(all='ignore'):
        for ia, a, ib, bf in _bz_cases(g):
            r = g['a%d_r' % ia]
            B1, B2, B3, om = ko.bfield_lab_omega(bf, a, r)
            ref = g['a%d_b%d' % (ia, ib)]
            for got, want in zip((B1, B2, B3), ref):
                assert_close(np.broadcast_to(got, r.shape), want, 1e-12, floor=1e-6 * np.max(np.abs(ref[np.isfinite(ref)])), what='B %s' % bf)
            assert_close(np.broadcast_to(om, r.shape), g['a%d_om%d' % (ia, ib)], 1e-12, floor=1e-9, what='omega %s' % bf)
            assert_close(ko.bfield_bsq(bf, vel, a, r), g['a%d_bsq%d' % (ia, ib)], 1e-11, what='bsq %s' % bf)
            for inu, nu in enumerate(g['nus']):
                u = np.array([np.broadcast_to(c, r.shape) for c in
                              ko.velocity_u_lab(dict(type='driftframe', bfield=bf, nu_parallel=float(nu)), a, r)])
                want = g['a%d_u%d_%d' % (ia, ib, inu)]
                # r = 2 with omega = 0 is the static limit: the drift speed is 1 +- eps there (gamma 1e8 or NaN)
                ok = g['a%d_ureal%d_%d' % (ia, ib, inu)] & ~(np.abs(want[0]) > 1e6)
                assert_close(u[:, ok], want[:, ok], 1e-11, floor=1.0, what='u %s nu=%g' % (bf, nu))


def test_bz_monopole_images_oracle(golden):
    g = golden('bz_models')
    cases = dict(mono=dict(type='bz_monopole'), power=dict(type='power', p=0.75))
    for name, bf in cases.items():
        a, th_o, mbar_max, half, npx, nu = [float(x) for x in g['img_%s_par' % name]]
        with np.errstate(all='ignore'):
            out = ko.make_image(a, np.inf, th_o, int(mbar_max), -half, half, -half, half, 2 * half / npx,
                                dict(type='ring', r_ring=3.5, sigma=0.8), dict(type='driftframe', bfield=bf, nu_parallel=nu), bf,
                                polarization=True, specind=1.)
        for got, key in zip(out, ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb']):
            ref = g['img_%s_%s' % (name, key)]
            if key in ('Q', 'U'):
                den = np.where(g['img_%s_I' % name] != 0, np.abs(g['img_%s_I' % name]), 1.0)
                assert np.nanmax(np.abs(got - ref) / den) <= 1e-10, (name, key)
            else:
                assert_close(got, ref, 1e-10, floor=1e-300, what='%s image %s' % (name, key))


def test_product_fr_table_matches_the_reference_recipe(golden):
    """kgeo_b200.bfields.bz_fr_table (NumPy dilogarithm; no SciPy in the product) vs the table the reference was run with.
    f(r) is a difference of terms ~r^3/4 that cancel to ~1/(4r): both sides carry ~r^3 eps of rounding noise."""
    from kgeo_b200.bfields import bz_fr_table
    g = golden('bz_models')
    rs, fs = bz_fr_table()
    assert rs.shape == (100000,) and rs[0] == 1.0 and rs[-1] == 100.0
    assert np.array_equal(rs[::250], g['fr_r'])
    d = np.abs(fs[::250] - g['fr_f'])
    assert np.all(d <= 1e-15 * g["fr_r"] ** 3 + 1e-15), float(np.max(d / (g['fr_r'] ** 3)))


# ---- field-line crossings of the jet models (kgeo/solver.py) -----------------------------------------
def _fieldline_case(g, name):
    a, r_o, th_o, psit, pval, neqmax, ngeo = [float(x) for x in g[name + '_par']]
    return (a, r_o, th_o, psit, pval, None if neqmax < 0 else int(neqmax), int(ngeo), str(g[name + '_model']),
            g[name + '_alpha'], g[name + '_beta'])


def test_fieldline_crossings_oracle_against_reference(golden):
    """getguesses / findroot / get_maxtau_forwardjet of the oracle (incl. SciPy's all-elements secant loop) vs the unmodified
    reference: identical guesses, and roots / coordinates / signs / crossing counts to 1e-12."""
    import warnings
    g = golden('fieldline')
    for name in json.loads(str(g['names'])):
        a, r_o, th_o, psit, pval, neqmax, ngeo, model, A, B = _fieldline_case(g, name)
        with np.errstate(all='ignore'), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            tau, _, (_, r, th, _) = ko.raytrace(a, [0, r_o, th_o, 0], [A, B], ngeo)
            guesses = ko.fieldline_guesses(r, th, tau, a, r_o, th_o, A, B, psit, neqmax, model, pval)
            res = ko.findroot(r, th, tau, psit, A, B, r_o, th_o, a, model, neqmax, 1e-8, pval)
            tm = ko.maxtau_forwardjet(a, r_o, th_o, A, B, neqmax)
        assert guesses.shape == g[name + '_guesses'].shape == tuple(g[name + '_shape'])
        assert_close(guesses, g[name + '_guesses'], 1e-13, what=name + ' guesses')
        assert_close(tm, g[name + '_taumax'], 1e-13, what=name + ' taumax')
        for k, v in zip(['tau', 'r', 'th', 'signpr', 'signpth', 'neq'], res[:6]):
            assert_close(v, g[name + '_' + k], 1e-12, what='%s %s' % (name, k))


def test_jet_pipeline_fixture_crossings(golden):
    """tests/golden/jet_pipeline.npz pins the targets of the rest of SURVEY 8 f4 (off_eq.getstokes of the unmodified reference:
    crossings, densities, per-crossing and sorted Stokes parameters), which is not built yet.  What IS built -- the crossings
    of solver.findroot -- is checked against it here, so the fixture and the oracle stay tied together."""
    import warnings
    g = golden('jet_pipeline')
    for name in json.loads(str(g['names'])):
        a, r_o, th_o, psit, pval, neqmax, ngeo = [float(x) for x in g[name + '_par']]
        A, B = g[name + '_alpha'].flatten(), g[name + '_beta'].flatten()
        with np.errstate(all='ignore'), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            tau, _, (_, r, th, _) = ko.raytrace(a, [0, r_o, th_o, 0], [A, B], int(ngeo))
            res = ko.findroot(r, th, tau, psit, A, B, r_o, th_o, a, name, int(neqmax), 1e-8, pval)
        assert tuple(res[6]) == tuple(g[name + '_shape'])
        for k, v in zip(['tau', 'r', 'th', 'signpr', 'signpth', 'neq'], res[:6]):
            assert_close(v, g[name + '_root_' + k], 1e-12, what='%s %s' % (name, k))
        n = len(g[name + '_density'])
        assert n == res[0].size and g[name + '_raw_i'].size == n and g[name + '_img_I'].shape[-2:] == g[name + '_alpha'].shape


def test_jet_pipeline_oracle_against_reference(golden):
    """The oracle's restatement of the whole off-equatorial jet pipeline -- crossings, stagnation surface (ff_boost), mass loading
    and density (densityfuncs), drift frame / field / redshift / Penrose-Walker transport off the equator (Iobs_off), subring
    sorting (geometry.sort_image) -- vs off_eq.getstokes of the unmodified reference for the monopole and power-law jets.
    This is the checker the device implementation of the rest of SURVEY 8 f4 will be held against."""
    import warnings
    g = golden('jet_pipeline')
    for name in json.loads(str(g['names'])):
        a, r_o, th_o, psit, pval, neqmax, ngeo = [float(x) for x in g[name + '_par']]
        A, B = g[name + '_alpha'], g[name + '_beta']
        shape = tuple(g[name + '_shape'])
        with np.errstate(all='ignore'), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            if name == 'mono':
                dens = ko.density_mono_all(g[name + '_root_r'], g[name + '_root_th'], shape, a, 0, 2)
            else:
                dens = ko.density_power_all(g[name + '_root_r'], g[name + '_root_th'], shape, psit, a, 0, 2, pval)
            raw = ko.getstokes(psit, A, B, r_o, th_o, a, int(ngeo), int(neqmax), 1e-8, name, pval, 0, 2, sumsubring=False)
            img = ko.getstokes(psit, A, B, r_o, th_o, a, int(ngeo), int(neqmax), 1e-8, name, pval, 0, 2)
        assert_close(dens, g[name + '_density'], 1e-12, what=name + ' density')
        for k, v in zip(['i', 'q', 'u', 'neq', 'r', 'th'], raw[:6]):
            ref = g[name + '_raw_' + k]
            assert_close(np.asarray(v, dtype=float), ref, 1e-11, floor=1e-6 * max(np.max(np.abs(ref)), 1e-300), what='%s raw %s' % (name, k))
        for k, v in zip(['I', 'Q', 'U', 'evpa', 'r', 'th', 'sinthb'], img):
            ref = g[name + '_img_' + k]
            assert_close(np.asarray(v, dtype=float), ref, 1e-11, floor=1e-6 * max(np.max(np.abs(ref)), 1e-300), what='%s image %s' % (name, k))


def test_jet_boost_oracle_against_reference(golden):
    """nu_parallel='FF' (velocities.getnu_cons with ff_boost.getEco) and the gammamax cap of u_driftframe in the oracle's getstokes
    vs the unmodified reference's (monopole jet; tests/golden/jet_boost.npz)."""
    import warnings
    g = golden('jet_boost')
    a, r_o, th_o, psit, neqmax, ngeo = [float(x) for x in g['par']]
    A, B = g['alpha'], g['beta']
    for name, kw in zip(json.loads(str(g['names'])), json.loads(str(g['kwargs']))):
        with np.errstate(all='ignore'), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            raw = ko.getstokes(psit, A, B, r_o, th_o, a, int(ngeo), int(neqmax), 1e-8, 'mono', 1, kw['nu_parallel'], 2, sumsubring=False,
                               gammamax=kw.get('gammamax'))
            img = ko.getstokes(psit, A, B, r_o, th_o, a, int(ngeo), int(neqmax), 1e-8, 'mono', 1, kw['nu_parallel'], 2,
                               gammamax=kw.get('gammamax'))
        for k, v in zip(['i', 'q', 'u', 'neq', 'r', 'th'], raw[:6]):
            ref = g[name + '_raw_' + k]
            assert_close(np.asarray(v, dtype=float), ref, 1e-11, floor=1e-6 * max(np.max(np.abs(ref)), 1e-300), what='%s raw %s' % (name, k))
        for k, v in zip(['I', 'Q', 'U', 'evpa', 'r', 'th', 'sinthb'], img):
            ref = g[name + '_img_' + k]
            assert_close(np.asarray(v, dtype=float), ref, 1e-11, floor=1e-6 * max(np.max(np.abs(ref)), 1e-300), what='%s image %s' % (name, k))


def test_integrals_oracle_against_reference(golden):
    """theta_motion / radial_motion / _S1_S2 / _R1_R2 / r_equatorial2 of the oracle at the interface of the reference's th_integrate,
    r_integrate, S1_S2, R1_R2, r_equatorial2 (called with the conserved quantities, incl. r_o = inf with r only)."""
    g = golden('integrals')
    for i in range(int(g['ncases'])):
        a, th_o, r_o = [float(x) for x in g['c%d_par' % i]]
        lam, eta, tau = g['c%d_lam' % i], g['c%d_eta' % i], g['c%d_tau' % i]
        r1, r2, r3, r4 = g['c%d_roots' % i]
        th = ko.theta_motion(a, th_o, g['c%d_so' % i], lam, eta, g['c%d_up' % i], g['c%d_um' % i], tau)
        rr = ko.radial_motion(a, r_o, lam, eta, r1, r2, r3, r4, tau)
        for got, ref, k in zip(th, g['c%d_th' % i], ('th', 'Gph', 'Gt')):
            assert_close(got, ref, 1e-12, floor=1e-3, what='case %d %s' % (i, k))
        for got, ref, k in zip(rr, g['c%d_r' % i], ('r', 'Iph', 'It', 'Isig')):
            assert_close(got, ref, 1e-12, floor=1e-3, what='case %d %s' % (i, k))
        assert np.all(g['c%d_zero' % i] == 0.0)          # do_phi_and_t=False leaves the other integrals zero ...
        assert np.array_equal(g['c%d_It_only' % i], -4 * tau)   # ... except I_t = 4 I_0 = -4 tau (:627-631)
        r_only = ko.radial_motion(a, r_o, lam, eta, r1, r2, r3, r4, tau, full=False)
        assert_close(r_only[0], g['c%d_r_only' % i], 1e-12, floor=1e-3, what='case %d r only' % i)
        assert np.array_equal(r_only[2], -4 * tau)
    r1, r2, r3, r4 = g['inf_roots']
    r_inf = ko.radial_motion(0.94, np.inf, g['inf_lam'], g['inf_eta'], r1, r2, r3, r4, g['inf_tau'], full=False)[0]
    assert_close(r_inf, g['inf_r'], 1e-12, floor=1e-3, what='r at r_o = inf')
    S1, S2 = ko._S1_S2(g['aux_al'], g['aux_phi'], g['aux_j'])
    R1, R2 = ko._R1_R2(g['aux_alR'], g['aux_phi'], g['aux_j'])
    for got, k in ((S1, 'aux_S1'), (S2, 'aux_S2'), (R1, 'aux_R1'), (R2, 'aux_R2')):
        assert_close(got, g[k], 1e-12, floor=1e-6, what=k)
    for (rho, vp, a_, t_, m), want in zip(g['obj_pts'], g['obj_r']):
        assert ko.r_equatorial2(rho, vp, a_, t_, mbar=int(m)) == pytest.approx(want, rel=1e-12)


def test_angular_turnings_oracle_against_reference(golden):
    g = golden('turnings')
    for i in range(int(g['ncases'])):
        a, th_o = [float(x) for x in g['c%d_par' % i]]
        A, B, tau = g['c%d_alpha' % i], g['c%d_beta' % i], g['c%d_tau' % i]
        assert_close(ko.n_angular_turnings(a, th_o, A, B, tau), g['c%d_n' % i], 1e-12, floor=1e-3, what='case %d' % i)
        assert_close(ko.n_angular_turnings(a, th_o, A, B, tau[3]), g['c%d_n1' % i], 1e-12, floor=1e-3, what='case %d, 1-D tau' % i)


def test_model_functions_oracle_against_reference(golden):
    """The oracle's closed forms against the reference's module-level u_* functions (velocities.py:179-470)."""
    g = golden('model_functions')
    for a in (0.0, 0.5, 0.94):
        tag = 'a%02d' % int(round(100 * a))
        r = g[tag + '_r']
        calls = dict(zamo=ko.u_zamo(a, r), infall=ko.u_infall(a, r), kep=ko.u_kep(a, r), kep_retro=ko.u_kep(a, r, retrograde=True),
                     subkep=ko.u_subkep(a, r, fac_subkep=0.7), gelles=ko.u_gelles(a, r, beta=0.4, chi=-2.0), simfit=ko.u_simfit(a, r),
                     general=ko.u_general(a, r, fac_subkep=0.7, beta_phi=0.6, beta_r=0.8))
        for k, u in calls.items():
            got = np.array([np.broadcast_to(np.asarray(c, dtype=float), r.shape) for c in u])
            assert_close(got, g[tag + '_u_' + k], 1e-12, floor=1e-12, what='%s %s' % (tag, k))
