"""Developer tool: ulp error of kg_math.cuh (host build) against mpmath-verified libm on random arguments."""
import ctypes, os, sys, numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
L = ctypes.CDLL(os.path.join(HERE, "_hostsim.so"))
dp = ctypes.POINTER(ctypes.c_double)
L.hostsim_math.argtypes = [ctypes.c_int, ctypes.c_size_t, dp, dp, dp]
def call(which, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float64); y = x if y is None else np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    L.hostsim_math(which, x.size, x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out
def ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))
rng = np.random.default_rng(7)
N = 4_000_000
np.seterr(all='ignore')
x = np.exp(rng.uniform(-700, 700, N)); print("log  wide    max ulp %.2f" % ulps(call(0, x), np.log(x)).max())
x = rng.uniform(0.5, 2.0, N); e = np.abs(call(0, x) - np.log(x)); print("log  [.5,2]  max ulp %.2f  max abs %.2e" % (ulps(call(0, x), np.log(x))[np.abs(x-1)>1e-3].max(), e.max()))
x = 1 + rng.uniform(-1e-3, 1e-3, N); print("log  near 1  max ulp %.2f" % ulps(call(0, x), np.log(x)).max())
for lim in (1.0, 100.0, 1e5):
    x = rng.uniform(-lim, lim, N)
    print("sin  |v|<%g max abs err %.2e   cos %.2e" % (lim, np.abs(call(1, x) - np.sin(x)).max(), np.abs(call(2, x) - np.cos(x)).max()))
x = rng.uniform(-0.8, 0.8, N); print("sin small max ulp %.2f cos %.2f" % (ulps(call(1, x), np.sin(x)).max(), ulps(call(2, x), np.cos(x)).max()))
y = rng.normal(size=N) * np.exp(rng.uniform(-20, 20, N)); x = rng.normal(size=N) * np.exp(rng.uniform(-20, 20, N))
print("atan2 max ulp %.2f  max abs %.2e" % (ulps(call(3, y, x), np.arctan2(y, x)).max(), np.abs(call(3, y, x) - np.arctan2(y, x)).max()))
y = rng.normal(size=N); x = rng.normal(size=N)
print("atan2 O(1) max ulp %.2f" % ulps(call(3, y, x), np.arctan2(y, x)).max())
x = rng.uniform(-1, 1, N); print("acos max ulp %.2f" % ulps(call(4, x), np.arccos(x)).max())
x = 1 - np.exp(rng.uniform(-36, -1, N)); print("acos near +1 max ulp %.2f   near -1 %.2f" % (ulps(call(4, x), np.arccos(x)).max(), ulps(call(4, -x), np.arccos(-x)).max()))
for v in (0.0, 1.0, -1.0, 0.5, -0.5): print("acos(%g) = %r vs %r" % (v, call(4, np.array([v]))[0], np.arccos(v)))
print("special: log(0)=%r log(inf)=%r log(-1)=%r atan2(0,1)=%r atan2(0,-1)=%r atan2(1,0)=%r acos(1.0000001)=%r" % (
    call(0, np.array([0.0]))[0], call(0, np.array([np.inf]))[0], call(0, np.array([-1.0]))[0], call(3, np.array([0.0]), np.array([1.0]))[0],
    call(3, np.array([0.0]), np.array([-1.0]))[0], call(3, np.array([1.0]), np.array([0.0]))[0], call(4, np.array([1.0000001]))[0]))

import mpmath as mp
mp.mp.dps = 40
for scale in (1.0, 1e-3, 1e-8):
    X = rng.normal(size=200000); Y = rng.normal(size=200000) * scale
    got = call(5, X, Y); ref = 0.5 * np.log(np.abs((X + Y) / (X - Y)))
    # the float64 reference itself loses eps * X / Y for small ratios: compare against atanh in extended precision there
    refx = np.array([float(mp.atanh(mp.mpf(float(y)) / mp.mpf(float(x)))) if abs(y) < abs(x) else float(mp.atanh(mp.mpf(float(x)) / mp.mpf(float(y)))) for x, y in zip(X[:20000], Y[:20000])])
    print("atanh_ratio scale %g: max abs err vs mp %.2e, max rel %.2e" % (scale, np.abs(got[:20000] - refx).max(), (np.abs(got[:20000] - refx) / np.abs(refx)).max()))
