"""Developer tool (GPU): the launches whose ncu captures are committed under profiles/ for the image / lensing kernels:
image_kernel unpolarized (default models) and polarized + pathlength at 256^2 (BASELINE cfg 2) and 4096^2, and
r_equatorial_kernel (n = 0..2) at 4096^2.  Prints CUDA-event times when run without a profiler."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgeo_b200 as kg
from kgeo_b200 import equatorial_images as ei, equatorial_lensing as el, _engine

a, th_o = 0.94, 17 * np.pi / 180
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for side in (256, 4096):
    x = np.linspace(-10, 10, side)
    A, B = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, x)]
    dA, dB = _engine.to_device(A), _engine.to_device(B)
    cases = [("unpol", lambda: ei.image_device(a, np.inf, th_o, 0, 2, dA, dB, ei.emis_default, ei.vel_default, ei.bfield_default,
                                                False, False, 1, 0.1, 230e9)),
             ("pol+path", lambda: ei.image_device(a, np.inf, th_o, 0, 2, dA, dB, ei.emis_default,
                                                  kg.Velocity('general', fac_subkep=0.7, beta_phi=0.7, beta_r=0.7),
                                                  kg.Bfield('bz_guess', C=1.), True, True, 1, 0.1, 230e9))]
    if side == 4096:
        cases.append(("r_equatorial", lambda: el._r_equatorial_device(a, np.inf, th_o, 0, 2, dA, dB)))
    for name, fn in cases:
        if not os.environ.get("PROFILE_ONCE"):
            fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("%5d^2 %-13s %8.3f ms  %.3e px/s" % (side, name, ms, side * side / ms * 1e3), flush=True)
