"""Developer tool: run the GPU track-parity cases through the host-compiled per-sample code and the tests' own gates."""
import sys, os, glob, numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from check import run
from helpers import track_errors, assert_tracks_close
from oracle import kgeo_oracle as ko
np.seterr(all='ignore')
force = int(sys.argv[1]) if len(sys.argv) > 1 else 0

def gate(got, ref, a, r_o, what, pixels=None):
    got = dict(got); got['tau'] = ref['tau']
    try:
        assert_tracks_close(got, dict(ref, a=a, r_o=r_o), what=what, pixels=pixels)
        what = "PASS " + what
    except AssertionError as e:
        what = "FAIL " + what + " (" + str(e)[:80] + ")"
    rplus = 1 + np.sqrt(1 - a * a)
    tau = ref['tau']
    worst = []
    for key, floor in (('r', 1e-3), ('th', 1.0), ('ph', 1.0), ('t', r_o), ('sig', r_o)):
        nanok = np.array_equal(np.isnan(got[key]), np.isnan(ref[key]))
        core, tail = track_errors(np.where(np.isnan(ref[key]), np.nan, got[key]), ref[key], tau, ref['r'], floor, rplus)
        worst.append("%s %.1e/%.0e%s" % (key, core, tail, "" if nanok else " NANPATTERN"))
        if core > 1e-10 or tail > 1e-2 or not nanok: worst[-1] += " <<<<<< FAIL"
    print("%-34s %s" % (what, "  ".join(worst)))

for f in sorted(glob.glob(os.path.join(ROOT, 'tests/golden/tracks_*.npz'))):
    g = np.load(f)
    a, th_o, r_o, ngeo = [float(x) for x in g['par']]
    (t, r, th, ph, sig), fl, w = run(a, r_o, th_o, g['alpha'], g['beta'], int(ngeo), force)
    gate(dict(t=t, r=r, th=th, ph=ph, sig=sig), {k: g[k] for k in ('tau', 'sig', 't', 'r', 'th', 'ph')}, a, r_o,
         os.path.basename(f) + " [%.0f%%/%.0f%%]" % (100 * np.mean((fl >> 14) & 1), 100 * np.mean((fl >> 15) & 1)),
         pixels=(th_o, g['alpha'], g['beta']))
rng = np.random.default_rng(1234)
for a, inc, r_o in ((0.94, 17.0, 1000.0), (0.5, 60.0, 500.0), (0.99, 80.0, 1000.0), (0.2, 163.0, 1000.0)):
    A = rng.uniform(-9, 9, 700); B = rng.uniform(-9, 9, 700)
    th_o = inc * np.pi / 180
    tau, sig, (t0, r0, th0, ph0) = ko.raytrace(a, [0, r_o, th_o, 0], [A, B], 40)
    (t, r, th, ph, sg), fl, w = run(a, r_o, th_o, A, B, 40, force)
    gate(dict(t=t, r=r, th=th, ph=ph, sig=sg), dict(tau=tau, sig=sig, t=t0, r=r0, th=th0, ph=ph0), a, r_o,
         "seeded a=%g inc=%g [%.0f%%/%.0f%%]" % (a, inc, 100 * np.mean((fl >> 14) & 1), 100 * np.mean((fl >> 15) & 1)),
         pixels=(th_o, A, B))
