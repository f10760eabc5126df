"""Developer tool (GPU): time kgeo_ray_setup + kgeo_tracks of BASELINE cfg 3 for every library build under
kgeo_b200/lib/variants/ in ONE process (torch is imported once), and compare each variant's r/theta/phi/t with the first one.

File names select the pass plan: lib_<name>@<plan>.so is loaded with KGEO_B200_TRK_PASSES=<plan> (the library reads the
variable once, on its first track launch).  Variants whose plan is polar / radial / fixup launch that pass alone (timing
only, no comparison)."""
import ctypes
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kgeo_b200 import _engine, _lib  # noqa: E402

SPIN, R_O, INC = 0.94, 1000.0, 17 * np.pi / 180
NSIDE = int(os.environ.get("KG_SWEEP_NSIDE", "1024"))
NGEO = int(os.environ.get("KG_SWEEP_NGEO", "500"))


def load(path):
    L = ctypes.CDLL(path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    x = np.linspace(-10, 10, NSIDE)
    A, B = [v.flatten() for v in np.meshgrid(x, x)]
    npix = A.size
    _lib._lib = load(sorted(glob.glob(os.path.join(ROOT, "kgeo_b200", "lib", "variants", "*.so")))[0])
    mode, _ = _engine.setup_mode(SPIN, INC, A, B, True)
    obs = _engine.observer_struct(SPIN, R_O, INC)
    d_alpha, d_beta = _engine.to_device(A, dev), _engine.to_device(B, dev)
    null = ctypes.cast(None, _lib.c_double_p)
    stream = _engine.stream_ptr()
    outs = [torch.empty((NGEO, npix), dtype=torch.float64, device=dev) for _ in range(4)]
    ref = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for so in sorted(glob.glob(os.path.join(ROOT, "kgeo_b200", "lib", "variants", "*.so"))):
        name = os.path.basename(so)[4:-3]
        plan = name.split("@")[1] if "@" in name else ""
        if plan:
            os.environ["KGEO_B200_TRK_PASSES"] = plan
        else:
            os.environ.pop("KGEO_B200_TRK_PASSES", None)
        L = load(so)
        nw = L.kgeo_ray_workspace_slots()
        d_w = torch.empty((int(L.kgeo_ray_workspace_doubles(npix)),), dtype=torch.float64, device=dev)
        d_flags = torch.empty((npix,), dtype=torch.int32, device=dev)
        for o in outs:
            o.fill_(float("nan"))

        def step():
            ev[0].record()
            rc = L.kgeo_ray_setup(ctypes.byref(obs), npix, _engine.dptr(d_alpha), _engine.dptr(d_beta), mode,
                                  _engine.dptr(d_w), _engine.iptr(d_flags), stream)
            ev[1].record()
            rc |= L.kgeo_tracks(ctypes.byref(obs), npix, _engine.dptr(d_w), _engine.iptr(d_flags), NGEO, 1 - 1e-10, null, 0,
                                1, _engine.dptr(outs[0]), _engine.dptr(outs[1]), _engine.dptr(outs[2]),
                                _engine.dptr(outs[3]), null, null, npix, stream)
            ev[2].record()
            if rc:
                raise RuntimeError(L.kgeo_last_error().decode())

        ts, tt = [], []
        for i in range(7):
            step()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(ev[0].elapsed_time(ev[1]))
                tt.append(ev[1].elapsed_time(ev[2]))
        msg = "%-34s setup %.3f ms  tracks %.3f ms (min %.3f)" % (name, np.median(ts), np.median(tt), min(tt))
        if plan not in ("polar", "radial", "fixup"):
            if ref is None:
                ref = [o.clone() for o in outs]
                msg += "  [reference for the comparison]"
            else:
                errs = []
                for o, r0, floor in zip(outs, ref, (R_O, 1.0, 1.0, 1.0)):
                    same_nan = bool((torch.isnan(o) == torch.isnan(r0)).all())
                    d = torch.nan_to_num((o - r0).abs() / torch.clamp(r0.abs(), min=floor), nan=0.0, posinf=0.0)
                    im = int(d.argmax())
                    errs.append("%.1e%s[k=%d pix=%d: %.17g vs %.17g; n(d>1e-12)=%d]" % (
                        float(d.max()), "" if same_nan else "(NaN pattern differs)", im // npix, im % npix,
                        float(o.view(-1)[im]), float(r0.view(-1)[im]), int((d > 1e-12).sum())))
                msg += "  max rel diff t,r,th,ph vs first: " + " ".join(errs)
        print(msg, flush=True)
        del d_w, d_flags


if __name__ == "__main__":
    main()
