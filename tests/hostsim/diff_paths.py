"""Developer tool: theta-function path vs Carlson path of the host-compiled per-sample code (no oracle)."""
import sys, numpy as np
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from check import run
a = float(sys.argv[1]); npx = int(sys.argv[2]); ngeo = int(sys.argv[3]); r_o = float(sys.argv[4]); th_o = float(sys.argv[5]) * np.pi / 180
lim = float(sys.argv[6]) if len(sys.argv) > 6 else 10.
x = np.linspace(-lim, lim, npx)
A, B = [v.flatten() for v in np.meshgrid(x, x)]
(t1, r1, th1, ph1, s1), fl1, w1 = run(a, r_o, th_o, A, B, ngeo, 1)
(t0, r0, th0, ph0, s0), fl0, w0 = run(a, r_o, th_o, A, B, ngeo, 0)
print("theta-polar %.2f%%  theta-radial %.2f%%" % (100 * np.mean((fl0 & (1 << 14)) != 0), 100 * np.mean((fl0 & (1 << 15)) != 0)))
cond = np.maximum(1.0, np.abs(r1) / 1000.)
for name, g, w, floor in (("r", r0, r1, 1.0), ("th", th0, th1, 1.0), ("ph", ph0, ph1, 1.0), ("t", t0, t1, r_o), ("sig", s0, s1, r_o)):
    e = np.abs(g - w) / np.maximum(np.abs(w), floor) / cond
    e[-1, :] = 0
    bad = np.isnan(g) != np.isnan(w)
    idx = np.unravel_index(np.nanargmax(e), e.shape)
    per_ray = np.nanmax(e, axis=0)
    print("%-3s max %.2e sample %d ray %d (alpha=%.5f beta=%.5f flags=%x) rays>1e-11: %d  >1e-10: %d  nanmismatch %d" % (
        name, np.nanmax(e), idx[0], idx[1], A[idx[1]], B[idx[1]], fl0[idx[1]], (per_ray > 1e-11).sum(), (per_ray > 1e-10).sum(), bad.sum()))
