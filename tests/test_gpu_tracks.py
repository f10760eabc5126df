"""GPU parity: per-ray setup and geodesic tracks through the public drop-in API (which calls the C ABI) vs
(1) fixtures produced by the unmodified reference, (2) the CPU oracle on seeded inputs, (3) size-independent invariants
at the BASELINE sizes, (4) the reference's own ipole known-answer rays."""
import numpy as np
import pytest

from helpers import PV_RTOL, assert_close, assert_tracks_close, removable_singularity_mask

pytestmark = pytest.mark.gpu
np.seterr(all='ignore')

TRACK_CASES = ['default', 'grid17', 'classII', 'extremal', 'spin0', 'south', 'midspin']


@pytest.fixture(scope="module")
def kg():
    import kgeo_b200
    return kgeo_b200


def _as_dict(geos):
    t, r, th, ph = geos.geo_coords
    return dict(tau=geos.mino_times, sig=geos.affine_times, t=t, r=r, th=th, ph=ph)


def test_setup_against_reference(golden, kg):
    from kgeo_b200 import _engine
    g = golden('setup')
    for i in range(int(g['ncases'])):
        a, th_o, r_o = [float(x) for x in g['c%d_par' % i]]
        A, B = g['c%d_alpha' % i], g['c%d_beta' % i]
        out = {k: v.cpu().numpy() for k, v in _engine.ray_summary(
            a, r_o, th_o, A, B, want=("roots", "rclass", "tau_tot", "uplus", "uminus", "uratio", "a2uminus", "outside")).items()}
        for key, gk in (("uplus", "up"), ("uminus", "um"), ("uratio", "ratio"), ("a2uminus", "a2um")):
            assert_close(out[key], g['c%d_%s' % (i, gk)], 1e-13, what='case %d %s' % (i, key))
        assert np.array_equal(out["rclass"], g['c%d_rclass' % i]), "root class differs in case %d" % i
        roots = g['c%d_roots' % i]
        scale = np.max(np.abs(roots), axis=0)
        for k in range(4):   # Ferrari's resolvent is ill-conditioned near double roots (class II / critical curve)
            assert np.max(np.abs(out["roots"][k] - roots[k].real) / scale) < 1e-9
            assert np.max(np.abs(out["roots"][4 + k] - roots[k].imag) / scale) < 1e-9
            assert np.array_equal(out["roots"][4 + k] == 0, roots[k].imag == 0)
        # the reference's own tau_tot is off by up to 4.6e-11 (vs 40-digit quadrature) where r2 -> 0 in these cases
        assert_close(out["tau_tot"], g['c%d_tmax' % i], 2e-10, what='tau_tot %d' % i)
        assert np.array_equal(out["outside"], g['c%d_outside' % i])
        lam = -A * np.sin(th_o)
        eta = (A ** 2 - a ** 2) * np.cos(th_o) ** 2 + B ** 2
        r1, r2, r3, r4, rclass = kg.radial_roots(a, lam, eta)
        assert np.array_equal(rclass, g['c%d_rclass' % i])
        assert_close(kg.mino_total(a, r_o, eta, *roots), g['c%d_tmax' % i], 2e-10, what='mino_total %d' % i)
        assert_close(kg.nmax_equatorial(a, r_o, th_o, A, B), g['c%d_nmax_eq' % i], 0.0, what='nmax_eq %d' % i)
        assert_close(kg.nmax_poloidal(a, r_o, th_o, A, B), g['c%d_nmax_pol' % i], 1e-10, what='nmax_pol %d' % i)
        up, um, thp, thm, thclass = kg.angular_turning(a, th_o, lam, eta)
        assert_close(up, g['c%d_up' % i], 1e-13)
        assert_close(um, g['c%d_um' % i], 1e-13)
        # theta_+- (kerr_raytracing_utils.py:257-262) against the unmodified reference; acos near u -> 1 has condition
        # number 1/sqrt(1 - u), so the gate is absolute on the angle
        assert np.max(np.abs(thp - g['c%d_thp' % i])) < 1e-12, "theta_plus differs in case %d" % i
        assert np.max(np.abs(thm - g['c%d_thm' % i])) < 1e-12, "theta_minus differs in case %d" % i
        assert np.array_equal(thclass, g['c%d_thclass' % i])


@pytest.mark.parametrize('name', TRACK_CASES)
def test_tracks_against_reference(golden, kg, name):
    g = golden('tracks_' + name)
    a, th_o, r_o, ngeo = [float(x) for x in g['par']]
    geos = kg.raytrace_ana(a, [0, r_o, th_o, 0], [g['alpha'], g['beta']], ngeo=int(ngeo), plotdata=False)
    ref = {k: g[k] for k in ('tau', 'sig', 't', 'r', 'th', 'ph')}
    ref.update(a=a, r_o=r_o)
    assert_tracks_close(_as_dict(geos), ref, what=name, pixels=(th_o, g['alpha'], g['beta']))
    assert geos.npix == len(g['alpha']) and geos.tautot.shape == (len(g['alpha']),)


def test_tracks_against_oracle_seeded(kg):
    """Same seeded random pixels through the CUDA path and the oracle (sizes the oracle finishes in seconds)."""
    from oracle import kgeo_oracle as ko
    rng = np.random.default_rng(1234)
    for a, inc, r_o in ((0.94, 17.0, 1000.0), (0.5, 60.0, 500.0), (0.99, 80.0, 1000.0), (0.2, 163.0, 1000.0)):
        A = rng.uniform(-9, 9, 700)
        B = rng.uniform(-9, 9, 700)
        th_o = inc * np.pi / 180
        tau, sig, (t, r, th, ph) = ko.raytrace(a, [0, r_o, th_o, 0], [A, B], 40)
        geos = kg.raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo=40, plotdata=False)
        ref = dict(tau=tau, sig=sig, t=t, r=r, th=th, ph=ph, a=a, r_o=r_o)
        assert_tracks_close(_as_dict(geos), ref, what='seeded a=%g inc=%g' % (a, inc), pixels=(th_o, A, B))


def test_coords_at_tau_matches_raytrace_and_partial_mode(golden, kg):
    g = golden('tracks_grid17')
    a, th_o, r_o, ngeo = [float(x) for x in g['par']]
    obs = [0, r_o, th_o, 0]
    sig, (t, r, th, ph) = kg.coords_at_tau(a, obs, [g['alpha'], g['beta']], g['tau'])
    ref = {k: g[k] for k in ('tau', 'sig', 't', 'r', 'th', 'ph')}
    ref.update(a=a, r_o=r_o)
    assert_tracks_close(dict(tau=g['tau'], sig=sig, t=t, r=r, th=th, ph=ph), ref, what='coords_at_tau',
                        pixels=(th_o, g['alpha'], g['beta']))
    # do_phi_and_t=False: r, theta identical, the rest zero (kerr_raytracing_ana.py:389-400)
    sig0, (t0, r0, th0, ph0) = kg.coords_at_tau(a, obs, [g['alpha'], g['beta']], g['tau'], do_phi_and_t=False)
    assert np.array_equal(r0, r) and np.array_equal(th0, th)
    assert not t0.any() and not ph0.any() and not sig0.any()
    # 1-D tau is one Mino time per ray
    s1, (t1, r1, th1, ph1) = kg.coords_at_tau(a, obs, [g['alpha'], g['beta']], g['tau'][3])
    assert r1.shape == (1, len(g['alpha'])) and np.array_equal(r1[0], r[3])


def test_ipole_midplane_intersections(kg):
    """The reference's tests/test_geodesics_kerr_ana.py:47-75 (all 15 rays, ngeo=10000, rtol 1e-3 vs ipole)."""
    tests = {
        (0.42, 60., -17., 0.): (-0.0522336, -16.0342), (0.42, 60., -9., 0.): (-0.103693, -8.09784),
        (0.42, 60., -4., 0.): (-0.308159, -3.20445), (0.42, 60., 1., 0.): (-0.327624, 1.67648),
        (0.42, 60., 6., 0.): (877.845, -13.6213), (0.42, 60., 8., 0.): (0.129720, 6.94829),
        (0.9, 10., -12., 0.): (-0.164094, -11.0082), (0.9, 10., -4., 0.): (-0.655582, -2.99558),
        (0.9, 10., 1., 0.): (-0.518830, -0.632352), (0.9, 10., 2., 0.): (-0.691096, -1.15767),
        (0.73, 35., 0.0001, -10): (11.8258, -0.0858775), (0.73, 35., 0.0001, -1.): (1.02201, 1.13478),
        (0.73, 35., 0.0001, 1.): (0.0148225, 0.0165869), (0.73, 35., 0.0001, 3.): (-1.20034, -1.08604),
        (0.73, 35., 0.0001, 4.): (-2.11221, 1.02322),
    }
    for (a, inc, al, be), want in tests.items():
        g = kg.raytrace_ana(a=a, observer_coords=[0, 1000., inc / 180. * np.pi, 0],
                            image_coords=[al, be], ngeo=10000, plotdata=False)   # scalars, as in the reference test
        _, r, th, ph = np.array(g.geo_coords)[:, :, 0]
        x, y, z = r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)
        if z[0] > z[1]:
            x, y, z = x[::-1], y[::-1], z[::-1]
        assert np.allclose([np.interp(0, z, x), np.interp(0, z, y)], want, rtol=1e-3), (a, inc, al, be)


def test_invariants_at_baseline_size(kg):
    """BASELINE cfg 3 geometry (a=0.94, 17 deg, r_o=1000, +-10 M) at full 1024^2 pixel count, few samples:
    size-independent properties -- r(0)=r_o, theta(0)=theta_o, phi(0)=t(0)=0, monotone tau, finite output, and
    theta = pi/2 to 1e-13 at the equatorial-crossing Mino times of r_equatorial (scripts/calc_t_equatorial.py:64-71)."""
    a, r_o, th_o = 0.94, 1000., 17 * np.pi / 180
    x = np.linspace(-10, 10, 1024)
    A, B = [v.flatten() for v in np.meshgrid(x, x)]
    geos = kg.raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo=6, plotdata=False)
    t, r, th, ph = geos.geo_coords
    assert r.shape == (6, 1024 * 1024)
    # the inversion tau -> r is ill-conditioned within ~1e-6 of the critical curve (k -> 1, amplitude -> pi/2): gate the
    # bulk tightly and the handful of pixels that straddle the curve loosely
    e0 = np.abs(r[0] / r_o - 1)
    assert np.mean(e0 > 1e-11) < 1e-3 and np.mean(e0 > 1e-9) < 1e-5 and np.max(e0) < 1e-5
    assert np.max(np.abs(th[0] - th_o)) < 1e-7 and np.mean(np.abs(th[0] - th_o) > 1e-12) < 1e-3
    assert np.mean(np.abs(ph[0]) > 1e-9) < 1e-4 and np.mean(np.abs(t[0]) > 1e-7) < 1e-4
    assert np.all(np.diff(geos.tausteps, axis=0) > 0)
    assert np.isfinite(r[:-1]).all() and np.isfinite(th).all()
    sub = slice(0, None, 37)
    rs, Ir, Imax, Nmax = kg.r_equatorial(a, r_o, th_o, 1, A[sub], B[sub])
    ok = (Nmax >= 1)
    sig, (tt, rr, thh, pp) = kg.coords_at_tau(a, [0, r_o, th_o, 0], [A[sub][ok], B[sub][ok]], Ir[ok])
    assert np.max(np.abs(thh[0] / (np.pi / 2) - 1)) < 1e-13
    assert np.max(np.abs(rr[0] / rs[ok] - 1)) < 1e-12


def test_theta_path_agrees_with_duplication_path_on_a_dense_grid(kg, monkeypatch):
    """The shipped per-ray theta-function evaluation vs the reference's algorithm (Carlson duplication per sample, selected
    with KGEO_B200_CARLSON_ONLY) over 512 x 512 rays x 40 samples of the BASELINE cfg 3 view and of an edge-on view: the
    two must agree to the parity gate on every sample that is not a removable singularity of the reference's formulas
    (see helpers.removable_singularity_mask), for all four coordinates."""
    from helpers import axis_weight, removable_singularity_mask, track_errors, TRACK_RTOL, PV_RTOL
    for a, inc, half in ((0.94, 17.0, 10.0), (0.7, 80.0, 8.0)):
        th_o, r_o = inc * np.pi / 180, 1000.
        x = np.linspace(-half, half, 512)
        A, B = [v.flatten() for v in np.meshgrid(x, x)]
        monkeypatch.delenv("KGEO_B200_CARLSON_ONLY", raising=False)
        g1 = kg.raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo=40, plotdata=False)
        monkeypatch.setenv("KGEO_B200_CARLSON_ONLY", "1")
        g0 = kg.raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo=40, plotdata=False)
        monkeypatch.delenv("KGEO_B200_CARLSON_ONLY", raising=False)
        assert np.array_equal(g0.tausteps, g1.tausteps)
        rplus = 1 + np.sqrt(1 - a * a)
        pv = removable_singularity_mask(a, r_o, th_o, A, B, g0.geo_coords[1])
        assert pv.mean() < 0.01
        axw = axis_weight(g0.geo_coords[2])      # phi within 0.01 rad of the rotation axis: arc length (oracle/gates.py)
        for key, got, ref, floor in (('r', g1.geo_coords[1], g0.geo_coords[1], 1e-3), ('th', g1.geo_coords[2], g0.geo_coords[2], 1.0),
                                     ('ph', g1.geo_coords[3], g0.geo_coords[3], 1.0), ('t', g1.geo_coords[0], g0.geo_coords[0], r_o)):
            wgt = axw if key == 'ph' else None
            if key in ('ph', 't'):
                core, _ = track_errors(np.where(pv, got, ref), ref, g0.tausteps, g0.geo_coords[1], floor, rplus, wgt)
                assert core <= PV_RTOL, "a=%g inc=%g %s on removable singularities: %.3e" % (a, inc, key, core)
                got = np.where(pv, ref, got)
            core, tail = track_errors(got, ref, g0.tausteps, g0.geo_coords[1], floor, rplus, wgt)
            assert core <= TRACK_RTOL, "a=%g inc=%g %s: theta vs duplication path %.3e" % (a, inc, key, core)
            assert tail <= 1e-2


@pytest.mark.parametrize('n,ngeo,half', [(203, 57, 9.0), (512, 41, 10.0)])
def test_two_launch_pass_equals_the_single_pass_kernel(kg, n, ngeo, half):
    """A full trace is a hot launch (theta-function code only, groups of four rays that are all hot) plus a fix-up launch (the
    other groups, complete code).  Against the one-pass kernel with the complete per-point code (full = 2): every ray is
    written exactly once (no NaN holes, no ray traced twice differently), r and theta are bit-identical, phi and t agree to the
    last bits away from the horizon (ptxas contracts the final sums differently under the hot kernel's register cap).
    203^2 rays x 57 samples: not a multiple of the 32-ray tile / the 64-sample pass; 512^2 x 41: the dense BASELINE cfg 3 view."""
    from kgeo_b200 import _engine
    from helpers import axis_weight, removable_singularity_mask, track_errors, PV_RTOL
    a, r_o, th_o = 0.94, 1000., 17 * np.pi / 180
    x = np.linspace(-half, half, n)
    A, B = [v.flatten() for v in np.meshgrid(x, x)]
    bundle = _engine.RayBundle(a, r_o, th_o, A, B)
    got = {k: v.cpu().numpy() for k, v in bundle.tracks(ntau=ngeo, maxtaufrac=1. - 1e-10, full=True).items()}
    ref = {k: v.cpu().numpy() for k, v in bundle.tracks(ntau=ngeo, maxtaufrac=1. - 1e-10, full=2).items()}
    assert np.array_equal(got['tau'], ref['tau'])
    assert np.array_equal(got['r'], ref['r'], equal_nan=True) and np.array_equal(got['th'], ref['th'], equal_nan=True)
    rplus = 1 + np.sqrt(1 - a * a)
    pv = removable_singularity_mask(a, r_o, th_o, A, B, ref['r'])
    for key, floor in (('ph', 1.0), ('t', r_o), ('sig', r_o)):
        assert np.array_equal(np.isnan(got[key]), np.isnan(ref[key])), key
        # Last-bit differences of the two instantiations are amplified (a) by 1 / sin(theta) in phi next to the rotation axis (a
        # coordinate singularity, gated on the arc length) and (b) by 1 / |1 - n sn^2| on the removable singularities of the
        # case-3/4 closed forms (gated at 1e-7 like everywhere else, oracle/gates.py)
        wgt = axis_weight(ref['th']) if key == 'ph' else None
        core, _ = track_errors(np.where(pv, got[key], ref[key]), ref[key], ref['tau'], ref['r'], floor, rplus, wgt)
        assert core <= PV_RTOL, (key, core)
        core, tail = track_errors(np.where(pv, ref[key], got[key]), ref[key], ref['tau'], ref['r'], floor, rplus, wgt)
        assert core <= 1e-11 and tail <= 1e-5, (key, core, tail)


def test_row_shards_equal_single_launch_bitwise(kg):
    """Multi-GPU check of SURVEY 8e runnable on one GPU: tracing row blocks separately == tracing the whole plane."""
    from kgeo_b200 import sharding
    a, r_o, th_o = 0.94, 1000., 17 * np.pi / 180
    x = np.linspace(-8, 8, 96)
    A, B = [v.flatten() for v in np.meshgrid(x, x)]
    whole = kg.raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo=20, plotdata=False)
    for world in (2, 3, 8):
        for rank in range(world):
            al, be, (r0, r1) = sharding.shard_pixels(A, B, 96, rank, world)
            part = kg.raytrace_ana(a, [0, r_o, th_o, 0], [np.ascontiguousarray(al), np.ascontiguousarray(be)], ngeo=20,
                                   plotdata=False)
            for w, p in zip(whole.geo_coords + [whole.mino_times, whole.affine_times],
                            part.geo_coords + [part.mino_times, part.affine_times]):
                assert np.array_equal(w[:, r0 * 96:r1 * 96], p, equal_nan=True)


def test_host_entry_point_matches_device_path(kg):
    """kgeo_raytrace_host (host buffers, slab pipeline) gives the same bits as the device-pointer path."""
    import ctypes
    from kgeo_b200 import _lib, _engine
    a, r_o, th_o = 0.94, 1000., 17 * np.pi / 180
    x = np.linspace(-8, 8, 64)
    A, B = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, x)]
    ngeo = 33
    geos = kg.raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo=ngeo, plotdata=False)
    outs = [np.empty((ngeo, A.size)) for _ in range(6)]
    mode, _ = _engine.setup_mode(a, th_o, A, B, True)
    obs = _engine.observer_struct(a, r_o, th_o)
    _lib.check(_lib.lib().kgeo_raytrace_host(ctypes.byref(obs), A.size, _engine.hptr(A), _engine.hptr(B), mode, ngeo,
                                             1. - 1.e-10, *[_engine.hptr(o) for o in outs], 0))
    t, r, th, ph = geos.geo_coords
    for got, ref in zip(outs, (t, r, th, ph, geos.affine_times, geos.mino_times)):
        assert np.array_equal(got, ref, equal_nan=True)


def test_argument_checks(kg):
    A = np.array([1.0, 2.0])
    with pytest.raises(Exception, match="should be a float in range"):
        kg.raytrace_ana(1, [0, 1000., 0.3, 0], [A, A], plotdata=False)
    with pytest.raises(Exception, match="r_o should be a float > 100"):
        kg.raytrace_ana(0.5, [0, 10., 0.3, 0], [A, A], plotdata=False)
    with pytest.raises(Exception, match="th_o should be a float"):
        kg.raytrace_ana(0.5, [0, 1000., np.pi / 2, 0], [A, A], plotdata=False)
    with pytest.raises(Exception, match="different lengths"):
        kg.raytrace_ana(0.5, [0, 1000., 0.3, 0], [A, np.array([1.0])], plotdata=False)
    with pytest.raises(Exception, match="incompatible shape"):
        kg.coords_at_tau(0.5, [0, 1000., 0.3, 0], [A, A], np.zeros((3, 5)))
    g = kg.raytrace_ana(0.5, [0, 1000., 0.3, 0], [3.0, 4.0], ngeo=5, plotdata=False)   # scalars are wrapped
    assert g.r_s.shape == (5, 1)


def test_edge_shapes_and_special_inputs(kg):
    """Ragged / degenerate shapes: empty pixel list, npix not a multiple of the 32-ray tile, ngeo = 1 and 2, one Mino
    time per ray, NaN pixel coordinates (NaN out, neighbours untouched), a sample count that is not a multiple of the warps."""
    obs = [0, 1000., 17 * np.pi / 180, 0]
    g0 = kg.raytrace_ana(0.94, obs, [np.zeros(0), np.zeros(0)], ngeo=5, plotdata=False)
    assert g0.r_s.shape == (5, 0)
    rng = np.random.default_rng(7)
    A, B = rng.uniform(-8, 8, 77), rng.uniform(-8, 8, 77)
    whole = kg.raytrace_ana(0.94, obs, [A, B], ngeo=23, plotdata=False)
    for n in (1, 31, 33, 77):
        part = kg.raytrace_ana(0.94, obs, [A[:n].copy(), B[:n].copy()], ngeo=23, plotdata=False)
        for w, p in zip(whole.geo_coords, part.geo_coords):
            assert np.array_equal(w[:, :n], p, equal_nan=True)
    one = kg.raytrace_ana(0.94, obs, [A, B], ngeo=1, plotdata=False)       # linspace(0, stop, 1) = [0]
    assert one.r_s.shape == (1, 77) and not one.mino_times.any()
    assert np.max(np.abs(one.r_s[0] / 1000. - 1)) < 1e-9
    two = kg.raytrace_ana(0.94, obs, [A, B], ngeo=2, plotdata=False)
    assert np.array_equal(two.mino_times[1], whole.mino_times[-1])         # last sample = maxtaufrac * tau_tot exactly
    A2, B2 = A.copy(), B.copy()
    A2[5] = np.nan
    bad = kg.raytrace_ana(0.94, obs, [A2, B2], ngeo=23, plotdata=False)
    assert np.isnan(bad.r_s[:, 5]).all()
    keep = np.arange(77) != 5
    assert np.array_equal(bad.r_s[:, keep], whole.r_s[:, keep], equal_nan=True)
    sig, (t, r, th, ph) = kg.coords_at_tau(0.94, obs, [A, B], np.full(77, 0.1))
    assert r.shape == (1, 77) and np.isfinite(r).all()


def test_more_than_2_pow_32_points_in_one_launch(kg):
    """BASELINE cfg 4 scale: one launch producing > 2^32 points (4096 x 2100 pixels x 500 samples = 4.3e9, a 34 GB array):
    64-bit indexing end to end.  The far corner of the big array must equal, bit for bit, a small launch on the same rays."""
    import torch
    from kgeo_b200 import _engine
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs 60 GB of free HBM")
    a, r_o, th_o = 0.94, 1000., 17 * np.pi / 180
    ncol, nrow, ngeo = 4096, 2100, 500
    x = np.linspace(-10, 10, ncol)
    y = np.linspace(-10, 10, nrow)
    A, B = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, y)]
    assert A.size * ngeo > 2 ** 32
    big = _engine.RayBundle(a, r_o, th_o, A, B)
    r = big.tracks(ntau=ngeo, maxtaufrac=1. - 1.e-10, want=("r",))["r"]
    assert tuple(r.shape) == (ngeo, A.size)
    head = r[0]
    assert float(torch.quantile((head[::97] / r_o - 1).abs(), 0.999)) < 1e-9
    tail_big = r[:, -64:].cpu().numpy()
    del r, big
    torch.cuda.empty_cache()
    small = _engine.RayBundle(a, r_o, th_o, A[-64:].copy(), B[-64:].copy())
    tail_small = small.tracks(ntau=ngeo, maxtaufrac=1. - 1.e-10, want=("r",))["r"].cpu().numpy()
    assert np.array_equal(tail_big, tail_small, equal_nan=True)
    assert np.isfinite(tail_big[:-1]).all()


def test_th_integrate_r_integrate_against_reference(golden):
    """th_integrate / r_integrate called with the conserved quantities, as the reference's solver and r_equatorial call them
    (kerr_raytracing_ana.py:195-365, 368-633), and S1_S2 / R1_R2 (:638-683) vs the unmodified reference."""
    from kgeo_b200 import kerr_raytracing_ana as ka
    g = golden('integrals')
    for i in range(int(g['ncases'])):
        a, th_o, r_o = [float(x) for x in g['c%d_par' % i]]
        lam, eta, tau = g['c%d_lam' % i], g['c%d_eta' % i], g['c%d_tau' % i]
        r1, r2, r3, r4 = g['c%d_roots' % i]
        th = ka.th_integrate(a, th_o, g['c%d_so' % i], lam, eta, g['c%d_up' % i], g['c%d_um' % i], tau)
        rr = ka.r_integrate(a, r_o, lam, eta, r1, r2, r3, r4, tau)
        ref_r = g['c%d_r' % i][0]
        cond = np.maximum(1.0, ref_r / 1000.)                     # the reference's own conditioning in r (oracle/gates.py)
        for got, ref, k, tol in zip(th, g['c%d_th' % i], ('th', 'Gph', 'Gt'), (1e-10, 1e-9, 1e-9)):
            assert got.shape == ref.shape
            assert_close(got, ref, tol, floor=1e-2, what='case %d %s' % (i, k))
        # the samples where the reference's own closed forms for the third-kind integrals are 0/0 expressions (oracle/gates.py):
        # gated at PV_RTOL there, as in the track tests
        alpha = -lam / np.sin(th_o)
        beta = g['c%d_so' % i] * np.sqrt(np.maximum(eta - (alpha ** 2 - a ** 2) * np.cos(th_o) ** 2, 0.0))
        pv = removable_singularity_mask(a, r_o, th_o, alpha, beta, ref_r)
        assert pv.mean() < 0.05                      # 2.8 % for the a = 0.998 case (12 x 12 rays within 3 M), 0.2-0.3 % otherwise
        for got, ref, k, tol in zip(rr, g['c%d_r' % i], ('r', 'Iph', 'It', 'Isig'), (1e-10, 1e-8, 1e-8, 1e-8)):
            floor = 1e-2 * (r_o if k in ('It', 'Isig') else 1.0)
            if k == 'r':
                assert_close(got / cond, ref / cond, tol, floor=floor, what='case %d %s' % (i, k))
                continue
            assert_close(np.where(pv, 0.0, got / cond), np.where(pv, 0.0, ref / cond), tol, floor=floor, what='case %d %s' % (i, k))
            assert_close(np.where(pv, got / cond, 0.0), np.where(pv, ref / cond, 0.0), PV_RTOL, floor=floor, what='case %d %s (pv)' % (i, k))
        th0 = ka.th_integrate(a, th_o, g['c%d_so' % i], lam, eta, g['c%d_up' % i], g['c%d_um' % i], tau, do_phi_and_t=False)
        rr0 = ka.r_integrate(a, r_o, lam, eta, r1, r2, r3, r4, tau, do_phi_and_t=False)
        assert_close(th0[0], g['c%d_th_only' % i], 1e-10, floor=1e-2, what='case %d th (r, theta only)' % i)
        assert_close(rr0[0] / cond, g['c%d_r_only' % i] / cond, 1e-10, floor=1e-2, what='case %d r (r, theta only)' % i)
        assert not th0[1].any() and not th0[2].any() and not rr0[1].any() and not rr0[3].any()
        assert np.array_equal(rr0[2], -4 * tau)
    r1, r2, r3, r4 = g['inf_roots']
    r_inf = ka.r_integrate(0.94, np.inf, g['inf_lam'], g['inf_eta'], r1, r2, r3, r4, g['inf_tau'], do_phi_and_t=False)[0]
    assert_close(r_inf, g['inf_r'], 1e-10, floor=1e-3, what='r at r_o = inf')
    S1, S2 = ka.S1_S2(g['aux_al'], g['aux_phi'], g['aux_j'])
    R1, R2 = ka.R1_R2(g['aux_alR'], g['aux_phi'], g['aux_j'])
    for got, k in ((S1, 'aux_S1'), (S2, 'aux_S2'), (R1, 'aux_R1'), (R2, 'aux_R2')):
        assert_close(got, g[k], 1e-10, floor=1e-6, what=k)
    with pytest.raises(Exception, match="not the same length"):
        ka.th_integrate(0.5, 0.3, np.ones(3), np.ones(2), np.ones(3), np.ones(3), np.ones(3), np.ones((2, 3)))
    with pytest.raises(Exception, match="incompatible shape in r_integrate"):
        ka.r_integrate(0.5, 1000., np.ones(3), np.ones(3), *([np.ones(3)] * 4), np.ones((2, 4)))
