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


def bz_fr_table():
    """(r, f(r)) of the reference's missing bz_fr_data.dat, from the formula in kgeo/bfields.py:7-13 (SciPy dilogarithm)."""
    from scipy.special import spence
    r = np.linspace(1.0, 100.0, 100000)
    x = 2.0 / r
    with np.errstate(all="ignore"):
        li = np.where(x <= 1, spence(1 - np.minimum(x, 1.0)),
                      np.pi ** 2 / 3.0 - 0.5 * np.log(x) ** 2 - spence(1 - np.minimum(1.0 / x, 1.0)))
        f = (li - np.log(np.abs(1 - x)) * np.log(r / 2.0)) * r * r * (2 * r - 3) / 8.0
        f += (1 + 3 * r - 6 * r * r) * np.log(r / 2.0) / 12.0
        f += 11.0 / 72.0 + 1.0 / (3.0 * r) + r / 2.0 - r * r / 2.0
    return r, f


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

    # (3) synthetic BZ-monopole table: the reference's own generating recipe (kgeo/bfields.py:7-20: f(r) on
    # linspace(1, 100, 100000), stored as 2 rows x N), with the real part of its complex mpmath formula written with
    # scipy.special.spence (Li2(x) = spence(1 - x); Re Li2(x) = pi^2/3 - ln^2(x)/2 - Li2(1/x) for x > 1).
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pkg_resources
    rs, fs = bz_fr_table()
    buf = io.StringIO()
    np.savetxt(buf, np.vstack((rs, fs)), fmt="%.18e")
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
