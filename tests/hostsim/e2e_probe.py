import os, sys, time, ctypes
import numpy as np, torch
sys.path.insert(0, os.getcwd())
os.environ["KGEO_B200_TRACE_HOST"] = "1"
from kgeo_b200 import _engine, _lib
L = _lib.lib()
th_o = 17 * np.pi / 180
x = np.linspace(-10, 10, 1024)
A, B = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, x)]
npix, NGEO = A.size, 500
mode, _ = _engine.setup_mode(0.94, th_o, A, B, True)
obs = _engine.observer_struct(0.94, 1000., th_o)
host = [torch.empty((NGEO, npix), dtype=torch.float64, pin_memory=True) for _ in range(4)]
hp = [ctypes.cast(h.data_ptr(), _lib.c_double_p) for h in host]
null = ctypes.cast(None, _lib.c_double_p)
for i in range(3):
    t0 = time.perf_counter()
    _lib.check(L.kgeo_raytrace_host(ctypes.byref(obs), npix, _engine.hptr(A), _engine.hptr(B), mode, NGEO, 1 - 1e-10, hp[0], hp[1], hp[2], hp[3], null, null, 0))
    print("call %d: %.1f ms total (incl. frees)" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
