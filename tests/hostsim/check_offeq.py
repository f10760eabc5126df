"""Developer tool (CPU): kg_offeq.cuh compiled for the host (hostsim_offeq.cpp) against the pinned oracle on the crossings of
tests/golden/jet_pipeline.npz: B^i, omega, drift-frame u^mu, redshift, sin(theta_b) and the per-crossing Stokes parameters.
Build first:  g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -Ikgeo_b200/csrc tests/hostsim/hostsim_offeq.cpp -o tests/hostsim/_hostsim_offeq.so"""
import ctypes, json, os, sys, warnings
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import kgeo_oracle as ko
warnings.simplefilter('ignore'); np.seterr(all='ignore')
worst = 0.0


def run(so_path, verbose=True):
    global worst
    L = ctypes.CDLL(so_path)
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda x: x.ctypes.data_as(dp)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'jet_pipeline.npz'))
    frx, frf = [np.ascontiguousarray(v) for v in ko.bz_fr_table()]
    worst = 0.0
    for name in json.loads(str(g['names'])):
        a, r_o, th_o, psit, pval, neqmax, ngeo = [float(x) for x in g[name + '_par']]
        shape = tuple(g[name + '_shape'])
        A, B = np.tile(g[name + '_alpha'].flatten(), shape[0]), np.tile(g[name + '_beta'].flatten(), shape[0])
        r, th = np.ascontiguousarray(g[name + '_root_r']), np.ascontiguousarray(g[name + '_root_th'])
        kr, kth = np.ascontiguousarray(g[name + '_root_signpr']), np.ascontiguousarray(g[name + '_root_signpth'])
        dens = np.ascontiguousarray(g[name + '_density'])
        n = r.size
        outs = [np.empty(n) for _ in range(5)]
        Bo, uo = np.empty(4 * n), np.empty(4 * n)
        L.hs_offeq(0 if name == 'mono' else 1, ctypes.c_double(1.0), ctypes.c_double(pval), 0, P(frx), P(frf), ctypes.c_longlong(frx.size),
                   ctypes.c_double(a), ctypes.c_double(th_o), n, P(A), P(B), P(r), P(th), P(kr), P(kth), ctypes.c_double(0.0), P(dens),
                   ctypes.c_double(1.0), *[P(o) for o in outs], P(Bo), P(uo))
        hit = r != 0
        bf = dict(type='bz_monopole') if name == 'mono' else dict(type='power', p=pval)
        vel = dict(type='driftframe', bfield=bf, nu_parallel=0)
        Bref = np.array([np.broadcast_to(x, r.shape) for x in ko.bfield_lab_omega(bf, a, r, th)])
        uref = np.array(ko.u_driftframe(a, r, bf, 0, th))
        I, Q, U, sinthb = ko.iobs_off(a, r_o, r, th_o, A, B, kr, kth, vel, bf, th, dens, 1)
        lam, eta = ko.conserved(a, th_o, A, B)
        gref = ko.redshift(a, r, lam, eta, kr, kth, *uref, th)

        def err(got, ref, what):
            global worst
            ok = hit & np.isfinite(ref)
            e = np.max(np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-6 * np.max(np.abs(ref[ok]))))
            worst = max(worst, e)
            assert np.array_equal(np.isnan(got[hit]), np.isnan(ref[hit])), what
            if verbose:
                print("%-6s %-8s max rel err %.2e" % (name, what, e))
        for i, k in enumerate(('B1', 'B2', 'B3', 'omega')):
            err(Bo[i::4], Bref[i], k)
        for i, k in enumerate(('u0', 'u1', 'u2', 'u3')):
            err(uo[i::4], uref[i], k)
        err(outs[4], gref, 'g')
        err(outs[3], np.broadcast_to(sinthb, r.shape), 'sinthb')
        # densities (jet_density: stagnation radius and normalisation from the product's host scalars)
        from kgeo_b200 import off_eq
        field, r0, norm = off_eq.jet_scalars(name, r, th, shape, psit, a, 2, pval)
        rho = np.empty(n)
        L.hs_jet_density(field, ctypes.c_double(1.0), ctypes.c_double(pval), 0, P(frx), P(frf), ctypes.c_longlong(frx.size),
                         ctypes.c_double(a), n, P(r), P(th), ctypes.c_double(0.0), ctypes.c_double(2.0), ctypes.c_double(r0),
                         ctypes.c_double(norm), ctypes.c_double(psit), P(rho))
        err(rho, dens, 'density')
        for got, ref, k in ((outs[0], I, 'I'), (outs[1], Q, 'Q'), (outs[2], U, 'U')):
            err(np.nan_to_num(got), np.nan_to_num(ref), k)      # getstokes applies nan_to_num to the Stokes parameters
            err(np.where(hit, np.nan_to_num(got), 0.0), g[name + '_raw_' + k.lower()], k + ' vs ref')

    return worst


if __name__ == "__main__":
    w = run(os.path.join(HERE, "_hostsim_offeq.so"))
    print("worst %.2e" % w)
    assert w < 1e-10
