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


# ---- parity gates of the GPU tests (SURVEY 8d "Parity gates") ----------------------------------------------------
TRACK_RTOL = 1.0e-10      # north_star: <= 1e-10 relative on r / theta / phi / t
IMAGE_RTOL = 1.0e-8       # north_star: images <= 1e-8 relative
HORIZON_TAIL = 0.999      # samples beyond this fraction of tau_tot on horizon-terminating rays are log-divergent


def track_errors(got, ref, tau, r_ref, floor, rplus):
    """Scaled error of one track array [ntau, npix] against the reference:  |got - ref| / (cond * max(|ref|, floor)).

    floor: phi crosses zero and theta is O(1) -> 1;  t and sigma are differences of antiderivatives of size ~r_o that
    cancel to 0 at tau = 0 (the reference itself returns t(0) ~ +-1e-9 instead of 0 at r_o = 1000) -> r_o;  r -> tiny.
    cond = max(1, r/1000): the map tau -> r has condition number ~ r / (r4 - r3), so next to a distant observer
    (r_o = 1e5 in the reference's default example) BOTH implementations are only good to ~1e-10 * r/1000 -- the exact value
    r(0) = r_o shows it: the reference returns 1e5 * (1 +- 2e-10) there.
    Returns (core, tail): max scaled error over well-conditioned samples, and over the log-divergent tail
    (tau > HORIZON_TAIL * tau_tot on rays that end on the horizon; kgeo/kerr_raytracing_ana.py:26 stops at
    (1 - 1e-10) tau_tot, where r - r_+ ~ 1e-10 and t, phi ~ log(r - r_+)).
    """
    got, ref = np.asarray(got), np.asarray(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN pattern differs"
    cond = np.maximum(1.0, np.abs(r_ref) / 1000.0)
    err = np.abs(got - ref) / (cond * np.maximum(np.abs(ref), floor))
    err = np.where(np.isfinite(err), err, 0.0)
    ends_on_horizon = np.nanmin(r_ref, axis=0) < rplus * (1 + 1e-3)
    tail = (tau > HORIZON_TAIL * tau[-1]) & ends_on_horizon[None, :]
    core = np.max(np.where(tail, 0.0, err)) if err.size else 0.0
    tl = np.max(np.where(tail, err, 0.0)) if err.size else 0.0
    return float(core), float(tl)


def assert_tracks_close(got, ref, what=""):
    """got/ref: dicts with tau, sig, t, r, th, ph [ntau, npix]; plus scalars a, r_o in ref."""
    a, r_o = float(ref['a']), float(ref['r_o'])
    rplus = 1 + np.sqrt(1 - a * a)
    tau = np.asarray(ref['tau'])
    # tau = linspace(0, tau_tot): inherits the reference's own root noise (its tau_tot is good to ~5e-11 against
    # 40-digit quadrature at high spin / small r2), hence the same 1e-10 gate
    assert_close(got['tau'], tau, TRACK_RTOL, what=what + ' tau')
    for key, floor in (('r', 1e-3), ('th', 1.0), ('ph', 1.0), ('t', r_o), ('sig', r_o)):
        core, tail = track_errors(got[key], ref[key], tau, np.asarray(ref['r']), floor, rplus)
        assert core <= TRACK_RTOL, "%s %s: scaled error %.3e > %.1e" % (what, key, core, TRACK_RTOL)
        assert tail <= 1e-2, "%s %s: horizon-tail error %.3e" % (what, key, tail)
