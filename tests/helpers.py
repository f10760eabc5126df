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


# ---- parity gates of the GPU tests (SURVEY 8d "Parity gates"): defined next to the oracle, oracle/gates.py ---------
from oracle.gates import (AXIS_SIN, HORIZON_TAIL, IMAGE_RTOL, PV_RTOL, PV_WIDTH, TRACK_RTOL,  # noqa: E402,F401
                          axis_weight, removable_singularity_mask, track_errors)


def assert_tracks_close(got, ref, what="", pixels=None):
    """got/ref: dicts with tau, sig, t, r, th, ph [ntau, npix]; plus scalars a, r_o in ref.
    pixels = (th_o, alpha, beta) enables the removable-singularity mask (see above) for phi, t, sigma."""
    a, r_o = float(ref['a']), float(ref['r_o'])
    rplus = 1 + np.sqrt(1 - a * a)
    tau = np.asarray(ref['tau'])
    pv = None
    if pixels is not None:
        pv = removable_singularity_mask(a, r_o, pixels[0], pixels[1], pixels[2], ref['r'])
        assert pv.mean() < 0.02, "%s: removable-singularity mask covers %.1f %% of the samples" % (what, 100 * pv.mean())
    # tau = linspace(0, tau_tot): inherits the reference's own root noise (its tau_tot is good to ~5e-11 against
    # 40-digit quadrature at high spin / small r2), hence the same 1e-10 gate
    assert_close(got['tau'], tau, TRACK_RTOL, what=what + ' tau')
    axw = axis_weight(ref['th'])          # phi within 0.01 rad of the rotation axis: gated on the arc length (oracle/gates.py)
    for key, floor in (('r', 1e-3), ('th', 1.0), ('ph', 1.0), ('t', r_o), ('sig', r_o)):
        g = np.asarray(got[key])
        wgt = axw if key == 'ph' else None
        if pv is not None and key in ('ph', 't', 'sig') and pv.any():
            core, _ = track_errors(np.where(pv, g, ref[key]), ref[key], tau, np.asarray(ref['r']), floor, rplus, wgt)
            assert core <= PV_RTOL, "%s %s: scaled error %.3e > %.1e on removable-singularity samples" % (what, key, core, PV_RTOL)
            g = np.where(pv, ref[key], g)
        core, tail = track_errors(g, ref[key], tau, np.asarray(ref['r']), floor, rplus, wgt)
        assert core <= TRACK_RTOL, "%s %s: scaled error %.3e > %.1e" % (what, key, core, TRACK_RTOL)
        assert tail <= 1e-2, "%s %s: horizon-tail error %.3e" % (what, key, tail)
