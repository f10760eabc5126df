"""Import shim that lets the UNMODIFIED reference (`/root/reference/kgeo`) run in the build container.

Only used by `tests/golden/make_golden.py` (fixture generation) and by optional
developer cross-checks.  Nothing on the GPU box imports this: `/root/reference`
does not exist there.  No reference source is copied; the package is imported in
place from its read-only mount.

The four shims are the ones listed in SURVEY.md section 8(c):
  1. stub modules for matplotlib / h5py / ehtim (imported at module top, never
     touched on the hot path with plotdata=False, savedata=False);
  2. np.infty / np.NaN aliases removed in NumPy 2;
  3. a synthetic `bz_fr_data.dat` served through `pkg_resources.resource_stream`
     (the blob is listed in the reference's `.MISSING_LARGE_BLOBS`; it only
     feeds Bfield('bz_monopole'), which is outside the hot path);
  4. the reference root on `sys.path`.
"""
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("KGEO_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave like a package so that submodule imports resolve
    sys.modules[name] = mod
    return mod


def install():
    """Install the shims and return the imported reference package."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "kgeo")):
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)

    # (2) NumPy-1 aliases
    if not hasattr(np, "infty"):
        np.infty = np.inf
    if not hasattr(np, "NaN"):
        np.NaN = np.nan

    # (1) plotting / IO / ehtim stubs
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "h5py", "ehtim",
                 "ehtim.parloop", "ehtim.observing", "ehtim.observing.obs_helpers",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["ehtim"].parloop = sys.modules["ehtim.parloop"]
    sys.modules["ehtim.observing"].obs_helpers = sys.modules["ehtim.observing.obs_helpers"]

    # (3) synthetic BZ-monopole table: 2 rows x N, r in [1,100]; the asymptotic
    # f(r) ~ 1/(4r) is good enough because no golden vector uses bz_monopole.
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pkg_resources
    rs = np.linspace(1.0, 100.0, 2000)
    buf = io.StringIO()
    np.savetxt(buf, np.vstack((rs, 0.25 / rs)))
    payload = buf.getvalue().encode()
    _orig_stream = pkg_resources.resource_stream

    def _resource_stream(package, resource):
        if resource == "bz_fr_data.dat":
            return io.BytesIO(payload)
        return _orig_stream(package, resource)

    pkg_resources.resource_stream = _resource_stream

    # (4) path
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import kgeo  # noqa: F401  (runs kgeo/__init__.py unchanged)
    import kgeo.scipy_ellip_binding  # noqa: F401
    import kgeo.velocities  # noqa: F401
    import kgeo.bfields  # noqa: F401
    import kgeo.emissivities  # noqa: F401
    return kgeo
