"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) in the build container.

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The GPU box has no /root/reference, so the fixtures are
committed; this script is committed with them so they can be regenerated and audited.
Each fixture stores its inputs next to the reference outputs.
SciPy/NumPy versions used are recorded in tests/golden/VERSIONS.json.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

kgeo = ref_shim.install()
import scipy  # noqa: E402
from kgeo.bfields import Bfield  # noqa: E402
from kgeo.emissivities import Emissivity  # noqa: E402
from kgeo.equatorial_images import make_image  # noqa: E402
from kgeo.equatorial_lensing import (critical_curve, nmax_equatorial, nmax_poloidal, r_equatorial,  # noqa: E402
                                     rho_of_req, rho_of_req_single)
from kgeo.kerr_raytracing_ana import coords_at_tau, raytrace_ana  # noqa: E402
from kgeo.kerr_raytracing_utils import (angular_turning, is_outside_crit, mino_total, n_equatorial_crossings,  # noqa: E402
                                         n_poloidal_orbits, radial_roots, uplus_uminus)
from kgeo.scipy_ellip_binding import ellip_pi_arr  # noqa: E402
from kgeo.velocities import Velocity  # noqa: E402
import scipy.special as sp  # noqa: E402

DEG = np.pi / 180.0


def quiet(fn, *args, **kwargs):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*args, **kwargs)


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-28s %8.1f kB' % (name, os.path.getsize(path) / 1e3))


def grid(n, half, n_beta=None, half_beta=None):
    al = np.linspace(-half, half, n)
    be = np.linspace(-(half_beta or half), (half_beta or half), n_beta or n)
    A, B = np.meshgrid(al, be)
    return A.flatten(), B.flatten()


# ---------------------------------------------------------------- special functions
def special():
    rng = np.random.default_rng(20261017)
    # the reference's own (print-only) check list: Pi(0.8; phi | -0.33), scipy_ellip_binding.py:241-256
    phis = np.array([.026, .1, np.pi / 2., 2.37, 4.1, 6., 7.7, 14.3, 215.])
    phis = np.concatenate([phis, -phis])
    ref_list = ellip_pi_arr(np.full(phis.shape, 0.8), phis, np.full(phis.shape, -0.33))
    # random sweep over the ranges the hot path visits incl. n>1 (rho<0 branch)
    N = 4000
    n = np.concatenate([rng.uniform(1e-6, 0.999, N // 2), rng.uniform(1.001, 80., N // 2)])
    m = np.concatenate([rng.uniform(-70., 0.999, N // 2), rng.uniform(0.001, 0.9999, N // 2)])
    phi = np.concatenate([rng.uniform(-19., 19., N // 2), rng.uniform(-np.pi, np.pi, N // 2)])
    pis = ellip_pi_arr(n, phi, m)
    # special-value masks
    n_sp = np.array([0.5, 0.5, 1.0, 2.0, np.inf, 0.3, 0.3, 4.0])
    phi_sp = np.array([5e-11, 1.0, 2.0, np.pi / 4, 1.0, np.pi / 2, -np.pi / 2, np.pi / 6])
    m_sp = np.array([0.3, -2e14, 0.5, 0.3, 0.2, 0.5, 0.5, 0.9])
    with np.errstate(all='ignore'):
        pis_sp = ellip_pi_arr(n_sp, phi_sp, m_sp)
    # Legendre F, E (cephes) incl. negative parameter; Jacobi; Carlson
    phi_fe = rng.uniform(-19., 19., N)
    m_fe = np.concatenate([rng.uniform(-70., 0.0, N // 2), rng.uniform(0.0, 0.99999, N // 2)])
    F = sp.ellipkinc(phi_fe, m_fe)
    E = sp.ellipeinc(phi_fe, m_fe)
    K = sp.ellipk(m_fe)
    u_j = rng.uniform(-20., 20., N)
    m_j = np.concatenate([rng.uniform(0.004, 0.9, N // 2), 1.0 - 10.0 ** rng.uniform(-9, -1, N // 2)])
    sn, cn, dn, am = sp.ellipj(u_j, m_j)
    x = 10.0 ** rng.uniform(-8, 1, N)
    y = 10.0 ** rng.uniform(-4, 2.5, N)
    z = 10.0 ** rng.uniform(-3, 1, N)
    p = 10.0 ** rng.uniform(-5, 1, N)
    x[:50] = 0.0
    save('special',
         pi_list_phi=phis, pi_list=ref_list,
         pi_n=n, pi_phi=phi, pi_m=m, pi=pis,
         pisp_n=n_sp, pisp_phi=phi_sp, pisp_m=m_sp, pisp=pis_sp,
         fe_phi=phi_fe, fe_m=m_fe, F=F, E=E, K=K,
         j_u=u_j, j_m=m_j, sn=sn, cn=cn, dn=dn, am=am,
         c_x=x, c_y=y, c_z=z, c_p=p,
         rf=sp.elliprf(x, y, z), rd=sp.elliprd(x, y, z), rj=sp.elliprj(x, y, z, p), rc=sp.elliprc(x, y))


# ---------------------------------------------------------------- per-ray constants
SETUP_CASES = [  # (a, th_o, r_o, half-width, n)
    (0.94, 17 * DEG, 1000., 8., 40),
    (0.94, 17 * DEG, np.inf, 10., 40),
    (0.5, 60 * DEG, 1.e5, 9., 32),
    (0.94, 80 * DEG, 1000., 3., 48),       # contains class-II pixels
    (0.99, 80 * DEG, np.inf, 3., 48),
    (0.998, 10 * DEG, 1000., 3., 32),      # class IV + vortical heavy
    (0.0, 45 * DEG, 1000., 8., 24),        # |a| < MINSPIN branches
    (0.7, 135 * DEG, 1000., 8., 24),       # th_o > pi/2
]


def setup():
    out = {}
    for i, (a, th_o, r_o, half, n) in enumerate(SETUP_CASES):
        A, B = grid(n, half)
        lam = -A * np.sin(th_o)
        eta = (A ** 2 - a ** 2) * np.cos(th_o) ** 2 + B ** 2
        up, um, ratio, a2um = uplus_uminus(a, th_o, lam.copy(), eta.copy())
        r1, r2, r3, r4, rclass = radial_roots(a, lam, eta)
        tmax = mino_total(a, r_o, eta, r1, r2, r3, r4)
        _, _, thp, thm, thclass = angular_turning(a, th_o, lam.copy(), eta.copy())      # theta_+-, kerr_raytracing_utils.py:242-274
        out.update({
            'c%d_thp' % i: thp, 'c%d_thm' % i: thm, 'c%d_thclass' % i: thclass,
            'c%d_par' % i: np.array([a, th_o, r_o]), 'c%d_alpha' % i: A, 'c%d_beta' % i: B,
            'c%d_up' % i: up, 'c%d_um' % i: um, 'c%d_ratio' % i: ratio, 'c%d_a2um' % i: a2um,
            'c%d_roots' % i: np.array([r1, r2, r3, r4]), 'c%d_rclass' % i: rclass, 'c%d_tmax' % i: tmax,
            'c%d_outside' % i: is_outside_crit(a, th_o, A, B),
            'c%d_nmax_eq' % i: quiet(nmax_equatorial, a, r_o, th_o, A, B),
            'c%d_nmax_pol' % i: quiet(nmax_poloidal, a, r_o, th_o, A, B),
        })
    out['ncases'] = np.array(len(SETUP_CASES))
    save('setup', **out)


# ---------------------------------------------------------------- tracks
TRACK_CASES = [  # (a, th_o, r_o, alpha, beta, ngeo)
    ('default', 0.94, 20 * DEG, 1.e5, np.linspace(-6, 6, 120), np.zeros(120), 25),   # BASELINE cfg 1, thinned
    ('grid17', 0.94, 17 * DEG, 1000., *grid(28, 8.), 24),
    ('classII', 0.94, 80 * DEG, 1000., *grid(30, 3.), 16),
    ('extremal', 0.998, 10 * DEG, 1000., *grid(20, 3.), 16),
    ('spin0', 0.0, 45 * DEG, 1000., *grid(16, 8.), 16),
    ('south', 0.7, 135 * DEG, 1000., *grid(16, 8.), 16),
    ('midspin', 0.42, 60 * DEG, 1000., *grid(16, 12.), 16),
]


def tracks():
    for (name, a, th_o, r_o, A, B, ngeo) in TRACK_CASES:
        with np.errstate(all='ignore'):
            g = quiet(raytrace_ana, a=a, observer_coords=[0, r_o, th_o, 0], image_coords=[A, B], ngeo=ngeo,
                      plotdata=False)
        save('tracks_' + name, par=np.array([a, th_o, r_o, ngeo]), alpha=A, beta=B,
             tau=g.mino_times, sig=g.affine_times,
             t=g.geo_coords[0], r=g.geo_coords[1], th=g.geo_coords[2], ph=g.geo_coords[3])


# ---------------------------------------------------------------- equatorial lensing / images
def lensing():
    out = {}
    cases = [(0.94, 17 * DEG, np.inf, 10., 48), (0.5, 60 * DEG, 1000., 9., 32), (0.0, 45 * DEG, np.inf, 8., 24),
             (0.7, 135 * DEG, 1000., 8., 24), (0.99, 85 * DEG, np.inf, 8., 32)]
    for i, (a, th_o, r_o, half, n) in enumerate(cases):
        A, B = grid(n, half)
        out['c%d_par' % i] = np.array([a, th_o, r_o])
        out['c%d_alpha' % i] = A
        out['c%d_beta' % i] = B
        for mbar in range(3):
            with np.errstate(all='ignore'):
                rs, Ir, Imax, Nmax = quiet(r_equatorial, a, r_o, th_o, mbar, A, B)
            out['c%d_m%d' % (i, mbar)] = np.array([rs, Ir, Imax, Nmax])
        # t, phi at the crossings: scripts/calc_t_equatorial.py:50-62 (needs finite r_o)
        if np.isfinite(r_o):
            _, Ir, _, _ = quiet(r_equatorial, a, r_o, th_o, 0, A, B)
            with np.errstate(all='ignore'):
                sig, co = quiet(coords_at_tau, a, [0, r_o, th_o, 0], [A, B], Ir, do_phi_and_t=True)
            out['c%d_cross0' % i] = np.array([sig[0], co[0][0], co[1][0], co[2][0], co[3][0]])
    out['ncases'] = np.array(len(cases))
    save('lensing', **out)


IMAGE_CASES = [
    # name, (a, r_o, th_o, mbar_max, half, npix), emissivity, velocity, bfield, kwargs
    ('default', (0.94, np.inf, 17 * DEG, 2, 10., 64), ('bpl', {}), ('zamo', {}), ('rad', {}), {}),
    ('testpol', (0.42, np.inf, 60 * DEG, 2, 10., 64), ('ring', {}),
     ('general', dict(fac_subkep=0.7, beta_phi=0.7, beta_r=0.7)), ('bz_guess', dict(C=1.)),
     dict(polarization=True, specind=1.)),
    ('subkep_pol', (0.7, np.inf, 45 * DEG, 2, 10., 48), ('ring', {}),
     ('general', dict(fac_subkep=0.6, beta_phi=0.8, beta_r=0.7)), ('bz_guess', dict(C=1.)),
     dict(polarization=True, specind=1.)),
    ('kep_vert_path', (0.9, 1000., 163 * DEG, 2, 10., 48), ('glm', {}), ('kep', {}), ('vert', {}),
     dict(polarization=True, pathlength=True, specind=0.5)),
    ('thermal_genpower', (0.9, np.inf, 30 * DEG, 1, 12., 40),
     ('thermal', dict(ne0=1.e5, Te0=1.e11, B0=5., bfield_model=True)), ('subkep', dict(fac_subkep=0.8)),
     ('gen_power', dict(n_I=0.5, p_val=1.0)), dict(polarization=True, pathlength=True)),
    ('infall_tor', (0.3, np.inf, 50 * DEG, 1, 10., 40), ('constant', {}), ('infall', {}), ('tor', {}), {}),
    ('gelles_simple', (0.6, np.inf, 40 * DEG, 1, 10., 40), ('bpl', dict(p1=-1.5, p2=-0.3, emiscut_in=3., emiscut_out=9.)),
     ('gelles', {}), ('simple', dict(Cr=0.5, Cvert=1.0, Cph=-0.7)), dict(polarization=True)),
    ('simfit_rm1', (0.8, np.inf, 25 * DEG, 1, 10., 40), ('ring', dict(r_ring=5., sigma=0.8)), ('simfit', {}),
     ('simple_rm1', dict(Cr=1.0, Cvert=0.2, Cph=0.3)), dict(polarization=True)),
    ('spin0_default', (0.0, np.inf, 45 * DEG, 2, 10., 40), ('bpl', {}), ('zamo', {}), ('rad', {}), {}),
]


def images():
    for (name, (a, r_o, th_o, mbar_max, half, npx), em, vel, bf, kw) in IMAGE_CASES:
        psize = 2 * half / npx
        with np.errstate(all='ignore'):
            out = quiet(make_image, a, r_o, th_o, mbar_max, -half, half, -half, half, psize,
                        emissivity=Emissivity(em[0], **em[1]), velocity=Velocity(vel[0], **vel[1]),
                        bfield=Bfield(bf[0], **bf[1]), **kw)
        names = ['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np', 'lp', 'Ie'][:len(out)]
        save('image_' + name, par=np.array([a, r_o, th_o, mbar_max, half, npx]),
             spec=np.array(json.dumps(dict(em=em, vel=vel, bf=bf, kw=kw))),
             **{k: v for k, v in zip(names, out)})


def models():
    """Element-wise physics closed forms on a radius grid."""
    r = np.linspace(1.05, 30., 160)
    out = {'r': r}
    spins = [0.0, 0.3, 0.94, -0.5]
    vels = [('zamo', {}), ('infall', {}), ('kep', {}), ('kep', dict(retrograde=True)), ('subkep', dict(fac_subkep=0.7)),
            ('general', dict(fac_subkep=0.6, beta_phi=0.8, beta_r=0.7)), ('general', dict(retrograde=True, fac_subkep=0.24, beta_phi=0.42, beta_r=0.24)),
            ('gelles', {}), ('simfit', {})]
    bfs = [('rad', {}), ('vert', {}), ('tor', {}), ('simple', dict(Cr=0.5, Cvert=1.0, Cph=-0.7)),
           ('simple_rm1', dict(Cr=1.0, Cvert=0.2, Cph=0.3)), ('bz_guess', dict(C=1.)), ('bz_guess', dict(C=-1.)),
           ('gen_power', dict(n_I=0.5, p_val=1.0)), ('gen_power', dict(n_I=0., p_val=0.75, isAbove=False)),
           ('monopoleA', dict(C=1.))]
    ems = [('constant', {}), ('bpl', {}), ('bpl', dict(p1=-1.5, p2=-0.3)), ('glm', {}), ('glm', dict(gamma_off=-2, sigma=1.)),
           ('ring', {}), ('ring', dict(r_ring=4, gamma_off=0.1, sigma=0.6))]
    out['spins'] = np.array(spins)
    out['spec'] = np.array(json.dumps(dict(vels=vels, bfs=bfs, ems=ems)))
    for ia, a in enumerate(spins):
        rr = r[r > 1 + np.sqrt(1 - a * a) + 0.02]
        out['a%d_r' % ia] = rr
        with np.errstate(all='ignore'):
            for iv, (vt, kw) in enumerate(vels):
                v = Velocity(vt, **kw)
                out['a%d_v%d' % (ia, iv)] = np.array([np.broadcast_to(c, rr.shape) for c in v.u_lab(float(a), rr)], dtype=float)
            for ib, (bt, kw) in enumerate(bfs):
                b = Bfield(bt, **kw)
                out['a%d_b%d' % (ia, ib)] = np.array([np.broadcast_to(c, rr.shape) for c in b.bfield_lab(float(a), rr)], dtype=float)
                out['a%d_bsq%d' % (ia, ib)] = b.bsq(float(a), rr, Velocity('general', fac_subkep=0.6, beta_phi=0.8, beta_r=0.7))
            for ie, (et, kw) in enumerate(ems):
                e = Emissivity(et, **kw)
                out['a%d_e%d' % (ia, ie)] = e.jrest(float(a), rr, g=0.7 + 0 * rr, sinthetab=0.6 + 0 * rr, specind=1.0)
            e = Emissivity('thermal', ne0=1.e5, Te0=1.e11, B0=5.)
            out['a%d_ethermal' % ia] = e.jrest(float(a), rr, g=0.7 + 0 * rr, sinthetab=0.6 + 0 * rr, Bmag=3.0 + 0 * rr)
    save('models', **out)


def counts():
    """n_equatorial_crossings / n_poloidal_orbits on full (ntau, npix) tau arrays."""
    a, th_o = 0.94, 17 * DEG
    A, B = grid(20, 8.)
    tau = np.linspace(0.0, 4.0, 9)[:, None] * np.ones(len(A))[None, :]
    save('counts', par=np.array([a, th_o]), alpha=A, beta=B, tau=tau,
         neq=n_equatorial_crossings(a, th_o, A, B, tau), npol=n_poloidal_orbits(a, th_o, A, B, tau))


def turnings():
    """n_angular_turnings (kerr_raytracing_utils.py:686-792) on (ntau, npix) and (npix,) tau arrays, ordinary and vortical rays."""
    from kgeo.kerr_raytracing_utils import n_angular_turnings
    out = {}
    cases = [(0.94, 17 * DEG, 8., 18), (0.998, 10 * DEG, 3., 16), (0.5, 90 * DEG, 6., 12)]
    for i, (a, th_o, half, n) in enumerate(cases):
        A, B = grid(n, half)
        tau = np.linspace(0.0, 4.0, 7)[:, None] * np.ones(len(A))[None, :]
        out.update({'c%d_par' % i: np.array([a, th_o]), 'c%d_alpha' % i: A, 'c%d_beta' % i: B, 'c%d_tau' % i: tau,
                    'c%d_n' % i: n_angular_turnings(a, th_o, A, B, tau), 'c%d_n1' % i: n_angular_turnings(a, th_o, A, B, tau[3])})
    out['ncases'] = np.array(len(cases))
    save('turnings', **out)


def model_functions():
    """The module-level closed forms of velocities.py (:179-470) and emissivities.py (:121-139) called directly."""
    from kgeo import velocities as kv, emissivities as ke
    out = {}
    r = np.concatenate([np.linspace(1.2, 3.0, 60), np.linspace(3.0, 60.0, 140)])
    out['r'] = r
    for a in (0.0, 0.5, 0.94):
        tag = 'a%02d' % int(round(100 * a))
        rr = r[r > 1 + np.sqrt(1 - a * a) + 1e-3]
        out[tag + '_r'] = rr
        calls = dict(zamo=kv.u_zamo(a, rr), infall=kv.u_infall(a, rr), kep=kv.u_kep(a, rr), kep_retro=kv.u_kep(a, rr, retrograde=True),
                     subkep=kv.u_subkep(a, rr, fac_subkep=0.7), gelles=kv.u_gelles(a, rr, beta=0.4, chi=-2.0),
                     simfit=kv.u_grmhd_fit(a, rr), general=kv.u_general(a, rr, fac_subkep=0.7, beta_phi=0.6, beta_r=0.8))
        for k, u in calls.items():
            out[tag + '_u_' + k] = np.array([np.broadcast_to(np.asarray(c, dtype=float), rr.shape) for c in u])
        out[tag + '_bpl'] = ke.emisBPL(a, rr)
        out[tag + '_bpl2'] = ke.emisBPL(a, rr, p1=-1.5, p2=-0.3)
        out[tag + '_glm'] = ke.emisGLM(a, rr)
        out[tag + '_ring'] = ke.emisGLM(a, rr, gamma_off=0.2, sigma=0.7, mu_ring=4.5)
        from kgeo import bfields as kb
        th = np.pi / 2
        bcalls = dict(simple=kb.Bfield_simple(a, rr, (0.87, 0.3, -0.5)), simple_rm1=kb.Bfield_simple_rm1(a, rr, (0.4, -0.7, 0.2)))
        if a > 0:
            bcalls.update(bzmagic=kb.Bfield_BZmagic(a, rr, th, C=-1.5), gen_power=kb.Bfield_gen_power(a, rr, th, 0.5, 0.8, True))
        for k, b in bcalls.items():
            out[tag + '_B_' + k] = np.array([np.broadcast_to(np.asarray(c, dtype=float), rr.shape) for c in b])
    save('model_functions', **out)


TABLE_FILE = '/root/reference/tests/data/a0p6_inflow.csv'      # committed copy: tests/golden/a0p6_inflow.csv


def fromfile_models():
    """Velocity('fromfile') / Bfield('fromfile') (np.interp tables; kgeo/velocities.py:626-698, kgeo/bfields.py:749-806):
    element-wise values on a radius grid that extends beyond both ends of the table, and one polarized image."""
    a = 0.6
    r = np.concatenate((np.linspace(1.81, 5.2, 150), [1.827254825200443022, 2.5, 6.0, 30.0]))
    v = Velocity('fromfile', file=TABLE_FILE)
    b = Bfield('fromfile', file=TABLE_FILE)
    out = dict(a=np.array(a), r=r)
    out['u'] = np.array([np.broadcast_to(c, r.shape) for c in v.u_lab(a, r)], dtype=float)
    out['B'] = np.array([np.broadcast_to(c, r.shape) for c in b.bfield_lab(a, r)], dtype=float)
    out['bsq'] = b.bsq(a, r, v)
    half, npx = 8., 40
    with np.errstate(all='ignore'):
        img = quiet(make_image, a, np.inf, 30 * DEG, 1, -half, half, -half, half, 2 * half / npx,
                    emissivity=Emissivity('ring', r_ring=3.5, sigma=0.6), velocity=v, bfield=b, polarization=True, specind=1.)
    out['par'] = np.array([a, np.inf, 30 * DEG, 1, half, npx])
    for k, val in zip(['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np'], img):
        out['img_' + k] = val
    save('fromfile_models', **out)


def bz_models():
    """Bfield('bz_monopole' | 'bz_para' | 'power') with omega_field, and Velocity('driftframe') on them (kgeo/bfields.py:
    545-709, kgeo/velocities.py:487-623): element-wise values and two polarized images.  The f(r) table the reference
    loads is the one tests/golden/ref_shim.py rebuilds; a decimated copy of what the reference interpolates is not needed
    because the fixture stores the table's generating grid parameters and the product rebuilds the same table."""
    from ref_shim import bz_fr_table
    rs, fs = bz_fr_table()
    out = dict(fr_r=rs[::250].copy(), fr_f=fs[::250].copy())
    bfs = [('bz_monopole', {}), ('bz_monopole', dict(C=-1.)), ('bz_monopole', dict(secondorder_only=True)), ('bz_para', {}),
           ('bz_para', dict(C=-1.)), ('power', dict(p=1)), ('power', dict(p=0.75, C=2.)), ('power', dict(p=0.75, usemono=True)),
           ('power', dict(p=0)), ('bz_guess', dict(C=1.)), ('monopoleA', dict(C=1.)), ('gen_power', dict(n_I=0.5, p_val=1.0))]
    nus = [0.0, 0.5, -0.3]
    spins = [0.3, 0.94, -0.5]
    out['spins'] = np.array(spins)
    out['nus'] = np.array(nus)
    out['spec'] = np.array(json.dumps(dict(bfs=bfs)))
    rgrid = np.concatenate((np.linspace(1.05, 12., 140), [1.9, 2.0, 2.5, 30., 99.5, 150.]))
    for ia, a in enumerate(spins):
        rp = 1 + np.sqrt(1 - a * a)
        rr = rgrid[rgrid > rp + 0.02]
        out['a%d_r' % ia] = rr
        with np.errstate(all='ignore'):
            for ib, (bt, kw) in enumerate(bfs):
                b = Bfield(bt, **kw)
                out['a%d_b%d' % (ia, ib)] = np.array([np.real(np.broadcast_to(c, rr.shape)) for c in b.bfield_lab(float(a), rr)], dtype=float)
                out['a%d_om%d' % (ia, ib)] = np.real(np.broadcast_to(b.omega_field(float(a), rr), rr.shape)).astype(float)
                out['a%d_bsq%d' % (ia, ib)] = np.real(b.bsq(float(a), rr, Velocity('general', fac_subkep=0.6, beta_phi=0.8, beta_r=0.7))).astype(float)
                for inu, nu in enumerate(nus):
                    v = Velocity('driftframe', bfield=b, nu_parallel=nu)
                    u = [np.broadcast_to(c, rr.shape) for c in v.u_lab(float(a), rr)]
                    out['a%d_u%d_%d' % (ia, ib, inu)] = np.array([np.real(c) for c in u], dtype=float)
                    # bz_para is complex-valued in the reference (scipy's lambertw): where E^2 > B^2 its square roots turn
                    # imaginary instead of NaN; record where the reference's result is genuinely real
                    out['a%d_ureal%d_%d' % (ia, ib, inu)] = np.all([np.imag(c) == 0 for c in u], axis=0)
    half, npx = 8., 36
    # (no bz_para image: scipy's lambertw makes that field complex-valued and the reference's Iobs cannot cast it)
    imgs = [('mono', 0.5, Bfield('bz_monopole'), 0.0), ('power', 0.7, Bfield('power', p=0.75), 0.3)]
    for name, a, b, nu in imgs:
        with np.errstate(all='ignore'):
            img = quiet(make_image, a, np.inf, 30 * DEG, 1, -half, half, -half, half, 2 * half / npx,
                        emissivity=Emissivity('ring', r_ring=3.5, sigma=0.8), velocity=Velocity('driftframe', bfield=b, nu_parallel=nu),
                        bfield=b, polarization=True, specind=1.)
        out['img_%s_par' % name] = np.array([a, 30 * DEG, 1, half, npx, nu])
        for k, val in zip(['I', 'Q', 'U', 'r', 't', 'g', 'sinthb', 'n', 'np'], img):
            out['img_%s_%s' % (name, k)] = np.real(val).astype(float)
    save('bz_models', **out)


def fieldline_crossings():
    """solver.getguesses / solver.findroot / solver.get_maxtau_forwardjet (kgeo/solver.py:57-207) on small ray grids: the three
    stream functions, neqmax 1 / 2 / None, a face-on distant-ray case that takes the flat-space guesses."""
    from kgeo.solver import findroot, getguesses, get_maxtau_forwardjet
    cases = [('para_n1', 0.5, 1000., 20 * DEG, grid(13, 9.), 2.0, 'para', 1, 1, 120),
             ('para_n2', 0.94, 1000., 60 * DEG, grid(11, 7.), 1.5, 'para', 1, 2, 150),
             ('mono_nall', 0.7, 500., 40 * DEG, grid(11, 8.), 0.3, 'mono', 1, None, 120),
             ('power_n1', 0.3, 1000., 163 * DEG, grid(11, 8.), 1.2, 'power', 0.75, 1, 120),
             ('para_flat', 0.5, 1.e4, 3 * DEG, grid(7, 300.), 40.0, 'para', 1, 1, 100)]
    out = dict(names=np.array(json.dumps([c[0] for c in cases])))
    for name, a, r_o, th_o, (A, B), psit, model, pval, neqmax, ngeo in cases:
        with np.errstate(all='ignore'):
            geo = quiet(raytrace_ana, a, [0, r_o, th_o, 0], [A, B], ngeo, plotdata=False)
            g = quiet(getguesses, geo, a, r_o, th_o, A, B, psit, ngeo, neqmax=neqmax, model=model, pval=pval)
            res = quiet(findroot, geo, psit, A, B, r_o, th_o, a, ngeo, model=model, neqmax=neqmax, tol=1e-8, pval=pval)
            tm = get_maxtau_forwardjet(a, r_o, th_o, A, B, neqmax=neqmax)
        out[name + '_par'] = np.array([a, r_o, th_o, psit, pval, -1 if neqmax is None else neqmax, ngeo])
        out[name + '_model'] = np.array(model)
        out[name + '_alpha'], out[name + '_beta'] = A, B
        out[name + '_guesses'] = g
        out[name + '_taumax'] = tm
        for k, v in zip(['tau', 'r', 'th', 'signpr', 'signpth', 'neq'], res[:6]):
            out[name + '_' + k] = np.asarray(v, dtype=float)
        out[name + '_shape'] = np.array(res[6])
    save('fieldline', **out)


def jet_pipeline():
    """Pinned targets for the rest of SURVEY 8 f4 (NOT built yet): off_eq.getstokes of the unmodified reference for the monopole and
    power-law jet models on small grids, with the intermediate quantities a device implementation has to reproduce: findroot's
    crossings, the densities of densityfuncs.density_*_all, the raw per-crossing Stokes parameters (sumsubring=False) and the
    sorted images.  The paraboloid model is absent: in this environment the reference's own pipeline fails for it
    (densityfuncs.eta_para: float() of a complex number -- SciPy's lambertw makes the field-line quantities complex)."""
    import contextlib
    import io
    from kgeo.off_eq import getstokes
    from kgeo.solver import findroot
    from kgeo.densityfuncs import density_mono_all, density_power_all
    out = dict(names=np.array(json.dumps(['mono', 'power'])))
    cases = [('mono', 0.5, 1000., 20 * DEG, 0.3, 1, 1), ('power', 0.7, 1000., 30 * DEG, 1.5, 0.75, 2)]
    for name, a, r_o, th_o, psit, pval, neqmax in cases:
        x = np.linspace(-8, 8, 11)
        A2, B2 = np.meshgrid(x, x)
        A, B = A2.flatten(), B2.flatten()
        ngeo = 100
        with np.errstate(all='ignore'), contextlib.redirect_stdout(io.StringIO()):
            geo = raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo, plotdata=False)
            fr = findroot(geo, psit, A, B, r_o, th_o, a, ngeo, model=name, neqmax=neqmax, tol=1e-8, pval=pval)
            if name == 'mono':
                dens = density_mono_all(fr[1], fr[2], fr[6], psit, a, 0, 2, neqmax=neqmax, gammamax=None)
            else:
                dens = density_power_all(fr[1], fr[2], fr[6], psit, a, 0, 2, neqmax=neqmax, gammamax=None, pval=pval)
            raw = getstokes(psit, A2, B2, r_o, th_o, a, ngeo, neqmax=neqmax, outgeo=geo, model=name, pval=pval, nu_parallel=0,
                            sigma=2, sumsubring=False)
            img = getstokes(psit, A2, B2, r_o, th_o, a, ngeo, neqmax=neqmax, outgeo=geo, model=name, pval=pval, nu_parallel=0,
                            sigma=2, retvals=True, retsin=True)
        out[name + '_par'] = np.array([a, r_o, th_o, psit, pval, neqmax, ngeo])
        out[name + '_alpha'], out[name + '_beta'] = A2, B2
        for k, v in zip(['tau', 'r', 'th', 'signpr', 'signpth', 'neq'], fr[:6]):
            out[name + '_root_' + k] = np.asarray(v, dtype=float)
        out[name + '_shape'] = np.array(fr[6])
        out[name + '_density'] = np.asarray(dens, dtype=float)
        for k, v in zip(['i', 'q', 'u', 'neq', 'r', 'th'], raw[:6]):
            out[name + '_raw_' + k] = np.real(np.asarray(v)).astype(float)
        for k, v in zip(['I', 'Q', 'U', 'evpa', 'r', 'th', 'sinthb'], img):
            out[name + '_img_' + k] = np.real(np.asarray(v)).astype(float)
    save('jet_pipeline', **out)


def jet_boost():
    """off_eq.getstokes of the unmodified reference for the monopole jet with the energy-conserving parallel boost
    (nu_parallel='FF': velocities.getnu_cons, ff_boost.getEco) and with a Lorentz-factor cap (gammamax): raw per-crossing Stokes
    parameters and the sorted images, on the grid of jet_pipeline's monopole case (same crossings)."""
    import contextlib
    import io
    from kgeo.off_eq import getstokes
    out = {}
    a, r_o, th_o, psit, neqmax, ngeo = 0.5, 1000., 20 * DEG, 0.3, 1, 100
    x = np.linspace(-8, 8, 11)
    A2, B2 = np.meshgrid(x, x)
    with np.errstate(all='ignore'), contextlib.redirect_stdout(io.StringIO()):
        geo = raytrace_ana(a, [0, r_o, th_o, 0], [A2.flatten(), B2.flatten()], ngeo, plotdata=False)
    cases = [('ff', dict(nu_parallel='FF')), ('ff_gmax', dict(nu_parallel='FF', gammamax=5.0)), ('gmax', dict(nu_parallel=0.3, gammamax=3.0))]
    out['par'] = np.array([a, r_o, th_o, psit, neqmax, ngeo])
    out['alpha'], out['beta'] = A2, B2
    out['names'] = np.array(json.dumps([c[0] for c in cases]))
    out['kwargs'] = np.array(json.dumps([c[1] for c in cases]))
    for name, kw in cases:
        with np.errstate(all='ignore'), contextlib.redirect_stdout(io.StringIO()):
            raw = getstokes(psit, A2, B2, r_o, th_o, a, ngeo, neqmax=neqmax, outgeo=geo, model='mono', sigma=2, sumsubring=False, **kw)
            img = getstokes(psit, A2, B2, r_o, th_o, a, ngeo, neqmax=neqmax, outgeo=geo, model='mono', sigma=2, retvals=True,
                            retsin=True, **kw)
        for k, v in zip(['i', 'q', 'u', 'neq', 'r', 'th'], raw[:6]):
            out[name + '_raw_' + k] = np.real(np.asarray(v)).astype(float)
        for k, v in zip(['I', 'Q', 'U', 'evpa', 'r', 'th', 'sinthb'], img):
            out[name + '_img_' + k] = np.real(np.asarray(v)).astype(float)
    save('jet_boost', **out)


def inverse_lensing():
    """rho_of_req / rho_of_req_single / critical_curve (equatorial_lensing.py:226-298)."""
    out = {}
    cases = [(0.94, 17 * DEG, 6.0), (0.5, 60 * DEG, 4.0), (0.99, 80 * DEG, 3.0)]
    varphis = np.linspace(-180, 165, 24) * np.pi / 180.
    out['cases'] = np.array(cases)
    out['varphis'] = varphis
    for ic, (a, th0, req) in enumerate(cases):
        for mbar in (0, 1, 2):
            vp, rho, al, be = quiet(rho_of_req, a, th0, req, mbar=mbar, varphis=varphis)
            out['rho_%d_%d' % (ic, mbar)] = rho
            out['alpha_%d_%d' % (ic, mbar)] = al
            out['beta_%d_%d' % (ic, mbar)] = be
    out['single'] = np.array([rho_of_req_single(0.94, 17 * DEG, 5.0, 0.3, mbar=m) for m in (0, 1, 2)])
    for ic, (a, th0, req) in enumerate(cases):
        al, be = critical_curve(a, th0, n=2001)
        out['crit_alpha_%d' % ic] = al
        out['crit_beta_%d' % ic] = be
    save('inverse_lensing', **out)


def integrals():
    """th_integrate / r_integrate called directly with the conserved quantities, as solver.py:168-195 and r_equatorial do
    (kerr_raytracing_ana.py:195-365, 368-633), their auxiliaries S1_S2 / R1_R2 (:638-683) and the Brent objective
    r_equatorial2 / objfunc (equatorial_lensing.py:198-224)."""
    from kgeo.kerr_raytracing_ana import R1_R2, S1_S2, r_integrate, th_integrate
    from kgeo.kerr_raytracing_utils import my_sign
    from kgeo.equatorial_lensing import objfunc, r_equatorial2
    out = {}
    cases = [(0.94, 17 * DEG, 1000., 8., 14), (0.998, 10 * DEG, 1000., 3., 12), (0.5, 120 * DEG, 500., 10., 12)]
    rng = np.random.default_rng(7)
    for i, (a, th_o, r_o, half, n) in enumerate(cases):
        A, B = grid(n, half)
        lam = -A * np.sin(th_o)
        eta = (A ** 2 - a ** 2) * np.cos(th_o) ** 2 + B ** 2
        s_o = my_sign(B)
        up, um, _, _ = uplus_uminus(a, th_o, lam.copy(), eta.copy())
        r1, r2, r3, r4, _ = radial_roots(a, lam, eta)
        tmax = mino_total(a, r_o, eta, r1, r2, r3, r4)
        tau = (np.linspace(0., 0.97, 9)[:, None] ** 1.5) * tmax[None, :] + 0 * rng.random((9, len(A)))
        th = th_integrate(a, th_o, s_o, lam, eta, up, um, tau)
        rr = r_integrate(a, r_o, lam, eta, r1, r2, r3, r4, tau)
        th0 = th_integrate(a, th_o, s_o, lam, eta, up, um, tau, do_phi_and_t=False)
        rr0 = r_integrate(a, r_o, lam, eta, r1, r2, r3, r4, tau, do_phi_and_t=False)
        out.update({'c%d_par' % i: np.array([a, th_o, r_o]), 'c%d_lam' % i: lam, 'c%d_eta' % i: eta, 'c%d_so' % i: s_o,
                    'c%d_up' % i: up, 'c%d_um' % i: um, 'c%d_roots' % i: np.array([r1, r2, r3, r4]), 'c%d_tau' % i: tau,
                    'c%d_th' % i: np.array(th), 'c%d_r' % i: np.array(rr), 'c%d_th_only' % i: np.array(th0[0]),
                    'c%d_r_only' % i: np.array(rr0[0]), 'c%d_It_only' % i: np.array(rr0[2]),      # I_t = 4 I_0 = -4 tau remains (:627-631)
                    'c%d_zero' % i: np.array([np.abs(th0[1]).max(), np.abs(th0[2]).max(), np.abs(rr0[1]).max(),
                                              np.abs(rr0[3]).max()])})
    # the observer at infinity, r only: what r_equatorial asks of r_integrate
    a, th_o = 0.94, 17 * DEG
    A, B = grid(12, 8.)
    lam = -A * np.sin(th_o)
    eta = (A ** 2 - a ** 2) * np.cos(th_o) ** 2 + B ** 2
    r1, r2, r3, r4, _ = radial_roots(a, lam, eta)
    tmax = mino_total(a, np.inf, eta, r1, r2, r3, r4)
    tau = np.linspace(0.05, 0.9, 6)[:, None] * tmax[None, :]
    out.update(inf_lam=lam, inf_eta=eta, inf_roots=np.array([r1, r2, r3, r4]), inf_tau=tau,
               inf_r=r_integrate(a, np.inf, lam, eta, r1, r2, r3, r4, tau, do_phi_and_t=False)[0])
    # auxiliaries at arguments of the ranges the radial cases produce
    al = rng.uniform(0.05, 3.0, 400)
    ph = rng.uniform(0.05, np.pi - 0.05, 400)
    j = rng.uniform(0.02, 0.98, 400)
    S1, S2 = S1_S2(al, ph, j)
    alR = rng.uniform(1.05, 6.0, 400)
    R1, R2 = R1_R2(alR, ph, j)
    out.update(aux_al=al, aux_phi=ph, aux_j=j, aux_S1=S1, aux_S2=S2, aux_alR=alR, aux_R1=R1, aux_R2=R2)
    # the Brent objective at a few image-plane points
    pts = [(5.5, 0.3, 0.94, 17 * DEG, 0), (5.2, -2.0, 0.94, 17 * DEG, 1), (2.0, 1.0, 0.94, 17 * DEG, 0), (30., 2.5, 0.5, 60 * DEG, 2),
           (4.0, 0.1, 0.5, 60 * DEG, 1)]
    out['obj_pts'] = np.array(pts)
    out['obj_r'] = np.array([quiet(r_equatorial2, rho, vp, a_, t_, mbar=int(m)) for rho, vp, a_, t_, m in pts])
    out['obj_f'] = np.array([quiet(objfunc, rho, vp, a_, t_, 4.0, mbar=int(m)) for rho, vp, a_, t_, m in pts])
    out['ncases'] = np.array(len(cases))
    save('integrals', **out)


if __name__ == '__main__':
    np.seterr(all='ignore')
    if len(sys.argv) > 1 and sys.argv[1] in ('integrals', 'turnings', 'model_functions', 'setup', 'inverse_lensing', 'fromfile_models', 'bz_models', 'fieldline_crossings', 'jet_pipeline', 'jet_boost'):     # add one fixture without touching the others
        globals()[sys.argv[1]]()
        sys.exit(0)
    special()
    setup()
    tracks()
    lensing()
    images()
    models()
    counts()
    inverse_lensing()
    fromfile_models()
    bz_models()
    fieldline_crossings()
    jet_pipeline()
    jet_boost()
    integrals()
    turnings()
    model_functions()
    with open(os.path.join(HERE, 'VERSIONS.json'), 'w') as f:
        json.dump(dict(numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0]), f)
