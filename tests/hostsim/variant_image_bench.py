"""Developer tool (GPU): time the fused image kernel (4096^2, n=0..2) for every library build under kgeo_b200/lib/variants/."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from kgeo_b200 import equatorial_images as ei, _engine
x = np.linspace(-10, 10, 4096); A, B = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, x)]
dA, dB = _engine.to_device(A), _engine.to_device(B)
for pol in (False, True):
    run = lambda: ei.image_device(0.94, np.inf, 17*np.pi/180, 0, 2, dA, dB, ei.emis_default, ei.vel_default, ei.bfield_default, pol, False, 1, 0.1, 230e9)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): r = run()
    e1.record(); torch.cuda.synchronize()
    print("pol=%%d %%.2f ms" %% (pol, e0.elapsed_time(e1)/5), end="  ")
print()
''' % ROOT
for so in sorted(glob.glob(os.path.join(ROOT, "kgeo_b200", "lib", "variants", "*.so"))):
    out = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, KGEO_B200_LIB=so), capture_output=True, text=True)
    print("%-24s %s" % (os.path.basename(so), out.stdout.strip() or out.stderr[-300:]), flush=True)
