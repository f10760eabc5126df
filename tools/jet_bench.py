"""GPU tool: timing of the off-equatorial jet pipeline (off_eq.getstokes: raytrace_ana + findroot + density + Iobs_off + sort) for
the monopole jet at 256 x 256 rays x 100 Mino-time samples, next to the oracle port of the reference on 32 x 32 rays.
Prints one JSON line; under ncu (`-k regex:fieldline|jet`) the same script gives the per-kernel times."""
import contextlib, io, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kgeo_b200 as kg
from kgeo_b200 import off_eq, _lib
np.seterr(all="ignore")
a, r_o, th_o, psit, ngeo = 0.5, 1000., 20 * np.pi / 180, 0.3, 100


def grid(n):
    x = np.linspace(-8, 8, n)
    return np.meshgrid(x, x)


def run_gpu(n, **kw):
    A2, B2 = grid(n)
    with contextlib.redirect_stdout(io.StringIO()):
        return off_eq.getstokes(psit, A2, B2, r_o, th_o, a, ngeo, neqmax=1, model='mono', sigma=2, **kw)


res = {}
for name, kw in (("nu0", dict(nu_parallel=0)), ("FF+gammamax", dict(nu_parallel='FF', gammamax=5.0))):
    run_gpu(256, **kw)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(3):
        out = run_gpu(256, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    res[name] = {"rays": 256 * 256, "ngeo": ngeo, "s_per_call": dt, "rays_per_s": 256 * 256 / dt,
                 "kernel_launches_per_call": (_lib.launch_count() - l0) / 3, "finite": bool(np.isfinite(out[0]).all())}
if not os.environ.get("KGEO_JET_NO_CPU"):
    from oracle import kgeo_oracle as ko
    import warnings
    warnings.simplefilter("ignore")
    A2, B2 = grid(32)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        ko.getstokes(psit, A2, B2, r_o, th_o, a, ngeo, 1, 1e-8, 'mono', 1, 0, 2)
    dt = time.perf_counter() - t0
    res["cpu_port_nu0"] = {"rays": 32 * 32, "ngeo": ngeo, "s_per_call": dt, "rays_per_s": 32 * 32 / dt, "cores": 1}
print(json.dumps(res))
