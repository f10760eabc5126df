"""GPU tool: the UNMASKED error budget of the track kernels against the CPU oracle (oracle/kgeo_oracle.py, pinned to the
unmodified reference by tests/golden) -> profiles/r02_parity_budget.json.

For every case: per coordinate the fraction of samples whose PLAIN relative error exceeds 1e-10, the maximum and the 99.9th
percentile, then the same with each stated relaxation of oracle/gates.py switched on one at a time (cond, floor, pv, tail,
axis) and cumulatively, and finally what the tests assert (`gated`).  Cases: BASELINE cfg 3 (4096 pixels spread over the
1024^2 grid, ngeo = 500), the four corners of the cfg-5 sweep (a = 0 / 0.99, theta_o = 5 / 85 deg; 2048 pixels of the 512^2
grid, ngeo = 200), and the reference's default example (r_o = 1e5: the `tracks_default` fixture's pixels, ngeo = 500, plus the
fixture itself, which is the unmodified reference's own output)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contextlib, io
import kgeo_b200 as kg
from oracle import kgeo_oracle as ko, gates
np.seterr(all="ignore")


def trace(a, r_o, th_o, A, B, ngeo):
    with contextlib.redirect_stdout(io.StringIO()):
        g = kg.raytrace_ana(a, [0, r_o, th_o, 0], [A, B], ngeo=ngeo, plotdata=False)
    t, r, th, ph = g.geo_coords
    return dict(t=np.array(t), r=np.array(r), th=np.array(th), ph=np.array(ph), tau=np.array(g.mino_times))


def case(name, a, r_o, th_o, A, B, ngeo, ref=None):
    t0 = time.time()
    if ref is None:
        tau, sig, (t, r, th, ph) = ko.raytrace(a, [0, r_o, th_o, 0], [A, B], ngeo)
        ref = dict(tau=tau, t=t, r=r, th=th, ph=ph)
    got = trace(a, r_o, th_o, A, B, ngeo)
    b = gates.parity_budget(got, ref, ref["tau"], a, r_o, th_o, A, B)
    b["case"] = name
    b["params"] = dict(a=a, r_o=r_o, th_o_deg=th_o * 180 / np.pi, pixels=int(A.size), ngeo=int(ngeo))
    b["tau_max_rel"] = float(np.nanmax(np.abs(got["tau"] - ref["tau"]) / np.maximum(np.abs(ref["tau"]), 1e-300)))
    b["passes_gates"] = gates.budget_passes(b)
    c = b["coords"]
    print("%-28s %6.1fs  pass=%s  plain frac>1e-10: r %.1e th %.1e ph %.1e t %.1e | gated max: r %.1e th %.1e ph %.1e t %.1e | "
          "pv %d tail %d axis %d of %d" % (name, time.time() - t0, b["passes_gates"], c["r"]["plain"]["frac_gt_1e10"],
          c["th"]["plain"]["frac_gt_1e10"], c["ph"]["plain"]["frac_gt_1e10"], c["t"]["plain"]["frac_gt_1e10"],
          c["r"]["gated"]["max"], c["th"]["gated"]["max"], c["ph"]["gated"]["max"], c["t"]["gated"]["max"],
          b["n_pv_masked"], b["n_tail"], b["n_axis"], b["n_samples"]), flush=True)
    return b


def spread(x, y, n):
    A, B = [v.flatten() for v in np.meshgrid(x, y)]
    idx = (np.arange(n, dtype=np.int64) * 257) % A.size
    return A[idx], B[idx]


out = []
x = np.linspace(-10, 10, 1024)
A, B = spread(x, x, 4096)
out.append(case("cfg3 (a=0.94, 17deg, r_o=1000)", 0.94, 1000., 17 * np.pi / 180, A, B, 500))
x5 = np.linspace(-10, 10, 512)
for a in (0.0, 0.99):
    for inc in (5.0, 85.0):
        A, B = spread(x5, x5, 2048)
        out.append(case("cfg5 corner a=%g inc=%g" % (a, inc), a, 1000., inc * np.pi / 180, A, B, 200))
g = np.load(os.path.join(ROOT, "tests", "golden", "tracks_default.npz"))
a, th_o, r_o, ngeo = [float(v) for v in g["par"]]
out.append(case("default example fixture (unmodified reference, r_o=1e5)", a, r_o, th_o, g["alpha"], g["beta"], int(ngeo),
                ref={k: g[k] for k in ("tau", "t", "r", "th", "ph")}))
out.append(case("default example, ngeo=500 (r_o=1e5)", a, r_o, th_o, g["alpha"], g["beta"], 500))
res = dict(what=__doc__, gates=dict(rtol=gates.TRACK_RTOL, pv_rtol=gates.PV_RTOL, pv_width=gates.PV_WIDTH, tail=gates.HORIZON_TAIL,
                                    tail_rtol=1e-2, axis_sin=gates.AXIS_SIN), cases=out)
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_parity_budget.json")
json.dump(res, open(dst, "w"), indent=1)
print("wrote", dst)
