"""Developer tool (GPU): small calls of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck).
Odd sizes on purpose: 41 x 37 rays (not a multiple of 4 or 32), 21 samples, plus a single-ray call."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kgeo_b200 as kg
a, th_o = 0.94, 17 * np.pi / 180
x = np.linspace(-8, 8, 41); y = np.linspace(-8, 8, 37)
A, B = [v.flatten() for v in np.meshgrid(x, y)]
g = kg.raytrace_ana(a, [0, 1000., th_o, 0], [A, B], ngeo=21, plotdata=False)
assert np.isfinite(g.geo_coords[1][:-1]).all()
g1 = kg.raytrace_ana(a, [0, 1000., th_o, 0], [A[:1], B[:1]], ngeo=7, plotdata=False)
sig, (t, r, th, ph) = kg.coords_at_tau(a, [0, 1000., th_o, 0], [A, B], g.tausteps[:5], do_phi_and_t=False)
img = kg.make_image(a, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 50, emissivity=kg.Emissivity('ring'),
                    velocity=kg.Velocity('general', fac_subkep=0.7, beta_phi=0.7, beta_r=0.7), bfield=kg.Bfield('bz_guess', C=1.),
                    polarization=True, specind=1., pathlength=True)
rs = kg.r_equatorial(a, np.inf, th_o, 1, A, B)
bm = kg.Bfield('bz_monopole')
img2 = kg.make_image(0.5, np.inf, 0.5, 1, -8., 8., -8., 8., 16. / 37, emissivity=kg.Emissivity('ring'),
                     velocity=kg.Velocity('driftframe', bfield=bm, nu_parallel=0.3), bfield=bm, polarization=True, specind=1.)
rr = np.linspace(1.9, 120., 333)
for bf in (bm, kg.Bfield('bz_para'), kg.Bfield('power', p=0.75)):
    bf.bfield_lab(0.5, rr); bf.omega_field(0.5, rr); kg.Velocity('driftframe', bfield=bf).u_lab(0.5, rr)
from kgeo_b200 import solver
gg = kg.raytrace_ana(0.5, [0, 1000., 20 * np.pi / 180, 0], [A[:333], B[:333]], ngeo=60, plotdata=False)
fr = solver.findroot(gg, 2.0, A[:333], B[:333], 1000., 20 * np.pi / 180, 0.5, 60, model='para', neqmax=2)
# jet pipeline (fieldline_* / jet_stokes / jet_sort kernels): odd sizes (7 x 9 rays, 47 samples), both jet models
from kgeo_b200 import off_eq
xa, xb = np.linspace(-7, 7, 7), np.linspace(-7, 7, 9)
A2, B2 = np.meshgrid(xa, xb)
for model, aa, inc, psit, pval, neqmax in (('mono', 0.5, 20., 0.3, 1, 1), ('power', 0.7, 30., 1.5, 0.75, 2)):
    st = off_eq.getstokes(psit, A2, B2, 1000., inc * np.pi / 180, aa, 47, neqmax=neqmax, model=model, pval=pval, nu_parallel=0, sigma=2,
                          retvals=True, retsin=True)
    assert np.isfinite(st[0]).all()
# round-2 additions: the integrals entry points, the curve library, the row-sharded image into a destination buffer (odd sizes) and
# the copy-engine placement
from kgeo_b200 import kerr_raytracing_ana as ka, kerr_raytracing_utils as ku, equatorial_lensing as el, sharding, equatorial_images as ei
lam, eta = -A * np.sin(th_o), (A ** 2 - a ** 2) * np.cos(th_o) ** 2 + B ** 2
up, um, _, _ = ku.uplus_uminus(a, th_o, lam, eta)
r1, r2, r3, r4, _ = ku.radial_roots(a, lam, eta)
tau5 = np.asarray(g.tausteps[:5])
ka.th_integrate(a, th_o, np.where(B >= 0, 1., -1.), lam, eta, up, um, tau5)
ka.r_integrate(a, 1000., lam, eta, r1, r2, r3, r4, tau5)
ka.r_integrate(a, np.inf, lam, eta, r1, r2, r3, r4, tau5, do_phi_and_t=False)
el.rho_library([0.3, 0.94, 0.7], [0.3, 1.0, 1.4], [3., 4., 5.], mmax=2, varphis=np.linspace(-3, 3, 13))
Ai, Bi = ei.pixel_grid(-10., 10., -10., 10., 20. / 37)
off, total = sharding.plane_layout(3, Ai.size)
buf = sharding.PeerBuffer(total)
for rank in range(3):
    al, be, rows = sharding.shard_pixels_interleaved(Ai, Bi, 37, rank, 3)
    ei.image_into(a, np.inf, th_o, 0, 2, al, be, *sharding._image_kwargs({}), dst={k: buf.ptr + 8 * o for k, o in off.items()},
                  cols=37, row_step=3, row_offset=rank, plane_stride=Ai.size)
flat = sharding.to_host_pinned(buf.tensor())
assert np.isfinite(flat[off["r"]:off["r"] + Ai.size]).all()
buf.close()
print("sanitize_case ok", g.geo_coords[1].shape, img[0].shape)
