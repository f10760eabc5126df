"""Near-minimax polynomial coefficients for the device elementary functions of kg_math.cuh (log, sincos, atan, asin).

Each function is written f(x) = x + x^3 R(x^2) (or cos: 1 - z/2 + z^2 C(z)); R is interpolated at Chebyshev nodes of the
z-interval in 60-digit arithmetic, converted to the monomial basis and rounded to double.  The script prints the C table and
the achieved maximum error of the ROUNDED polynomial (evaluated in 60 digits) relative to f, so the degree choice is
auditable.  Run: python research/gen_libm_coeffs.py
"""
import mpmath as mp
mp.mp.dps = 60


def cheb_fit(R, z0, z1, deg):
    n = deg + 1
    nodes = [(z0 + z1) / 2 + (z1 - z0) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, z in enumerate(nodes):
        for j in range(n):
            A[i, j] = z ** j
        b[i] = R(z)
    c = mp.lu_solve(A, b)
    return [float(c[j]) for j in range(n)]


def poly(c, z):
    s = mp.mpf(0)
    for a in reversed(c):
        s = s * z + mp.mpf(a)
    return s


def report(name, c, f_true, f_poly, xs):
    worst = max(abs((f_poly(x) - f_true(x)) / f_true(x)) for x in xs if f_true(x) != 0)
    print("// %s: degree %d in z, max rel err of the rounded polynomial %.2e" % (name, len(c) - 1, float(worst)))
    print("//   " + ", ".join("%.17e" % a for a in c))
    return c


def grid(a, b, n=400):
    return [a + (b - a) * mp.mpf(k + 0.5) / n for k in range(n)]


out = {}
# log: 2 atanh(s) = 2 s + s^3 R(s^2), |s| <= (sqrt2-1)/(sqrt2+1)
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
R = lambda z: (2 * mp.atanh(mp.sqrt(z)) - 2 * mp.sqrt(z)) / (z * mp.sqrt(z)) if z > 0 else mp.mpf(2) / 3
c = cheb_fit(R, mp.mpf(0), smax ** 2 * 1.02, 6)
out['log'] = report("log: 2atanh(s) = 2s + s^3 R(s^2)", c, lambda s: 2 * mp.atanh(s), lambda s: 2 * s + s ** 3 * poly(c, s * s), grid(mp.mpf(1e-4), smax))
# sin / cos on |r| <= pi/4 (a little wider)
zmax = (mp.pi / 4 * 1.01) ** 2
R = lambda z: (mp.sin(mp.sqrt(z)) - mp.sqrt(z)) / (z * mp.sqrt(z)) if z > 0 else -mp.mpf(1) / 6
c = cheb_fit(R, mp.mpf(0), zmax, 6)
out['sin'] = report("sin r = r + r^3 S(r^2)", c, mp.sin, lambda r: r + r ** 3 * poly(c, r * r), grid(mp.mpf(1e-4), mp.pi / 4))
R = lambda z: (mp.cos(mp.sqrt(z)) - 1 + z / 2) / (z * z) if z > 0 else mp.mpf(1) / 24
c = cheb_fit(R, mp.mpf(0), zmax, 5)
out['cos'] = report("cos r = 1 - z/2 + z^2 C(z)", c, mp.cos, lambda r: 1 - r * r / 2 + r ** 4 * poly(c, r * r), grid(mp.mpf(1e-4), mp.pi / 4))
# atan on |t| <= tan(pi/16)
tmax = mp.tan(mp.pi / 16) * 1.005
R = lambda z: (mp.atan(mp.sqrt(z)) - mp.sqrt(z)) / (z * mp.sqrt(z)) if z > 0 else -mp.mpf(1) / 3
c = cheb_fit(R, mp.mpf(0), tmax ** 2, 7)
out['atan'] = report("atan t = t + t^3 A(t^2)", c, mp.atan, lambda t: t + t ** 3 * poly(c, t * t), grid(mp.mpf(1e-4), tmax))
# asin on |x| <= 0.5
R = lambda z: (mp.asin(mp.sqrt(z)) - mp.sqrt(z)) / (z * mp.sqrt(z)) if z > 0 else mp.mpf(1) / 6
c = cheb_fit(R, mp.mpf(0), mp.mpf(0.2505), 12)
out['asin'] = report("asin x = x + x^3 Q(x^2)", c, mp.asin, lambda x: x + x ** 3 * poly(c, x * x), grid(mp.mpf(1e-4), mp.mpf(0.5)))
print()
flat = []
idx = {}
for k in ('log', 'sin', 'cos', 'atan', 'asin'):
    idx[k] = len(flat)
    flat += out[k]
print("// offsets:", idx, "total", len(flat))
print("#define KG_MATH_COEFFS {" + ", ".join("%.17e" % a for a in flat) + "}")
