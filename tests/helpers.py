"""Shared comparison helpers for the parity tests."""
import json

import numpy as np


def rel_err(got, ref, floor=0.0):
    """max |got-ref| / max(|ref|, floor), NaN/inf positions must coincide."""
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    bad_ref = ~np.isfinite(ref)
    bad_got = ~np.isfinite(got)
    assert np.array_equal(np.isnan(ref), np.isnan(got)), "NaN pattern differs"
    assert np.array_equal(ref[bad_ref & ~np.isnan(ref)], got[bad_ref & ~np.isnan(ref)]), "inf pattern differs"
    ok = ~(bad_ref | bad_got)
    if not ok.any():
        return 0.0
    den = np.maximum(np.abs(ref[ok]), floor) if floor > 0 else np.abs(ref[ok])
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(got[ok] - ref[ok]) / den))


def assert_close(got, ref, rtol, floor=0.0, what=""):
    e = rel_err(got, ref, floor)
    assert e <= rtol, "%s: max rel err %.3e > %.1e" % (what, e, rtol)


def image_models(spec_json):
    """Decode the model spec stored in an image fixture into oracle-style dicts."""
    spec = json.loads(str(spec_json))
    em = dict(type=spec['em'][0], **spec['em'][1])
    vel = dict(type=spec['vel'][0], **spec['vel'][1])
    bf = dict(type=spec['bf'][0], **spec['bf'][1])
    return em, vel, bf, spec['kw']
