"""Developer tool (GPU): kernel-only time of the fused image kernel (4096^2, n = 0..2; unpolarized default models and polarized +
path length) and of r_equatorial_kernel with preallocated outputs, for the default library and every build under
kgeo_b200/lib/variants/."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import ctypes, sys, numpy as np, torch
sys.path.insert(0, %r)
import kgeo_b200 as kg
from kgeo_b200 import equatorial_images as ei, _engine, _lib
from kgeo_b200.models import pack_models
L = _lib.lib(); dev = _engine.device(); null = ctypes.cast(None, _lib.c_double_p); st = _engine.stream_ptr()
side = 4096; npx = side * side
x = np.linspace(-10, 10, side); A, B = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, x)]
dA, dB = _engine.to_device(A), _engine.to_device(B)
img = torch.empty((30, npx), dtype=torch.float64, device=dev)
ptrs = [_engine.dptr(img[3 * i:3 * i + 3]) for i in range(7)] + [_engine.dptr(img[21 + i]) for i in range(3)]
obs = _engine.observer_struct(0.94, float("inf"), 17 * np.pi / 180)
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
res = []
for pol, path, vel, bf in ((0, 0, ei.vel_default, ei.bfield_default), (1, 1, kg.Velocity('general', fac_subkep=0.7, beta_phi=0.7, beta_r=0.7), kg.Bfield('bz_guess', C=1.))):
    M = pack_models(ei.emis_default, vel, bf)
    opts = _lib.ImageOpts(0, 2, pol, path, 1.0, 0.1, 230e9)
    extra = [_engine.dptr(img[24:27]), _engine.dptr(img[27:30])] if path else [null, null]
    res.append(timeit(lambda: _lib.check(L.kgeo_equatorial_image(ctypes.byref(obs), ctypes.byref(M), ctypes.byref(opts), npx, _engine.dptr(dA), _engine.dptr(dB), *ptrs, *extra, st))))
res.append(timeit(lambda: _lib.check(L.kgeo_r_equatorial(ctypes.byref(obs), npx, _engine.dptr(dA), _engine.dptr(dB), 0, 2, _engine.dptr(img[0:3]), _engine.dptr(img[3:6]), _engine.dptr(img[6:9]), _engine.dptr(img[9]), _engine.dptr(img[10]), st))))
print("image unpol %%.3f ms   pol+path %%.3f ms   r_equatorial %%.3f ms" %% tuple(res))
''' % ROOT
runs = [("default", {})] + [(os.path.basename(so), {"KGEO_B200_LIB": so}) for so in sorted(glob.glob(os.path.join(ROOT, "kgeo_b200", "lib", "variants", "*.so")))]
for name, extra in runs:
    out = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, **extra), capture_output=True, text=True)
    print("%-24s %s" % (name, out.stdout.strip() or out.stderr[-400:]), flush=True)
