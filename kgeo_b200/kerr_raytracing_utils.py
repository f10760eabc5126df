"""Drop-in for kgeo/kerr_raytracing_utils.py on the hot path: the Geodesics result type and the per-ray helper
functions, computed by libkgeo_b200.so kernels (no NumPy arithmetic on geodesic quantities here).

savegeos / loadgeos write / read the reference's HDF5 layout when h5py is installed (ImportError otherwise).
Out of scope: plotgeos (matplotlib host-side code) and n_angular_turnings.
"""
import numpy as np

from . import _checks, _engine

MINSPIN = 1.e-4  # kgeo/kerr_raytracing_utils.py:9
EP = 1.e-12      # kgeo/kerr_raytracing_utils.py:12


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else x


class Geodesics(object):
    """Result of raytrace_ana (kgeo/kerr_raytracing_utils.py:15-93): same fields and properties.

    The (ngeo, npix) arrays are produced on first use: the reference's NumPy-valued fields (`mino_times`, `affine_times`,
    `geo_coords`, `r_s`, ...) through the host pipeline of the C ABI (slabs of samples computed while the previous slab drains
    into page-locked memory), `device_array(name)` as device-resident [ngeo, npix] float64 torch tensors from one launch.  Both
    are cached; objects built from NumPy arrays (loadgeos, user code) just hold them.
    """

    def __init__(self, a, observer_coords, image_coords, mino_times, affine_times, geo_coords, _bundle=None,
                 _device_tracks=None, _host_tracks=None):
        self.a = a
        self.observer_coords = observer_coords
        self.image_coords = image_coords
        self._dev = dict(tau=mino_times, sig=affine_times)
        for name, arr in zip(("t", "r", "th", "ph"), geo_coords):
            self._dev[name] = arr
        self._host = {}
        self._bundle = _bundle
        self._device_tracks = _device_tracks      # names -> dict of [ngeo, npix] CUDA tensors (one kernel launch)
        self._host_tracks = _host_tracks          # names -> dict of (ngeo, npix) NumPy arrays (slab pipeline of the C ABI)

    # ---- lazy host views ----
    def device_array(self, name):
        """[ngeo, npix] CUDA tensor of 't','r','th','ph','sig','tau' (NumPy inputs are returned unchanged)."""
        if self._dev.get(name) is None:
            if self._device_tracks is None:
                raise RuntimeError("Geodesics: array %r was not computed" % name)
            # the four coordinates come out of one launch: ask for all that are still missing
            want = tuple(n for n in ("t", "r", "th", "ph") if self._dev.get(n) is None) if name in ("t", "r", "th", "ph") else (name,)
            for k, v in self._device_tracks(want).items():
                if self._dev.get(k) is None:
                    self._dev[k] = v
        return self._dev[name]

    def _materialise(self, names):
        """Bring the named arrays to the host in ONE overlapped pass: page-locked NumPy arrays (torch's caching pinned
        allocator: a block freed by an earlier result is reused without re-pinning), one asynchronous D2H copy per array
        queued back to back, one synchronisation.  A script that reads `geo_coords` sees the PCIe link rate instead of four
        serial pageable copies."""
        todo = [n for n in names if n not in self._host]
        if not todo:
            return
        if self._host_tracks is not None and all(self._dev.get(n) is None for n in todo):
            self._host.update(self._host_tracks(todo))    # nothing on the device yet: straight through the host pipeline
            return
        t = _engine.torch()
        pending = []
        for name in todo:
            src = self.device_array(name)
            if not hasattr(src, "is_cuda"):               # NumPy input (loadgeos, user-constructed objects)
                self._host[name] = src
                continue
            try:
                dst = t.empty(src.shape, dtype=src.dtype, pin_memory=True)
            except RuntimeError:                          # page-locked memory exhausted: pageable destination
                dst = t.empty(src.shape, dtype=src.dtype)
            dst.copy_(src, non_blocking=True)
            pending.append((name, dst, src.device))
        for dev in set(d for _, _, d in pending):
            t.cuda.current_stream(dev).synchronize()
        for name, dst, _ in pending:
            self._host[name] = dst.numpy()                # the array keeps the pinned block alive (arr.base)

    def _get(self, name):
        self._materialise((name,))
        return self._host[name]

    @property
    def mino_times(self):
        return self._get("tau")

    @property
    def affine_times(self):
        return self._get("sig")

    @property
    def geo_coords(self):
        self._materialise(("t", "r", "th", "ph"))
        return [self._host["t"], self._host["r"], self._host["th"], self._host["ph"]]

    # observer
    @property
    def t_o(self):
        return self.observer_coords[0]

    @property
    def r_o(self):
        return self.observer_coords[1]

    @property
    def th_o(self):
        return self.observer_coords[2]

    @property
    def ph_o(self):
        return self.observer_coords[3]

    # image
    @property
    def alpha(self):
        return self.image_coords[0]

    @property
    def beta(self):
        return self.image_coords[1]

    @property
    def npix(self):
        return len(self.alpha)

    @property
    def lam(self):
        return -self.alpha * np.sin(self.th_o)

    @property
    def eta(self):
        return (self.alpha ** 2 - self.a ** 2) * np.cos(self.th_o) ** 2 + self.beta ** 2

    @property
    def n_poloidal(self):  # fractional number of poloidal orbits
        return n_poloidal_orbits(self.a, self.th_o, self.alpha, self.beta, self.device_array("tau"))

    @property
    def nmax_eq(self):  # number of equatorial crossings
        return n_equatorial_crossings(self.a, self.th_o, self.alpha, self.beta, self.device_array("tau")[-1])

    # geodesics
    @property
    def tausteps(self):
        return self.mino_times

    @property
    def tautot(self):
        return self.mino_times[-1]

    @property
    def affinesteps(self):
        return self.affine_times

    @property
    def t_s(self):
        return self._get("t")

    @property
    def r_s(self):
        return self._get("r")

    @property
    def th_s(self):
        return self._get("th")

    @property
    def ph_s(self):
        return self._get("ph")

    @property
    def sig_s(self):
        return self.affine_times

    def plotgeos(self, *args, **kwargs):
        raise NotImplementedError("plotgeos is not built: host-side matplotlib code outside the accelerated path (SURVEY 2, out of scope)")

    def savegeos(self, outfile=None, path='./'):
        """HDF5 file with the reference's dataset names (kgeo/kerr_raytracing_utils.py:201-221).  Needs h5py."""
        h5py = _require_h5py()
        if outfile is None:
            fname = path + 'a%0.2f_th%0.2f_geo.h5' % (self.a, self.th_o * 180 / np.pi)
        else:
            fname = outfile
        self._materialise(("t", "r", "th", "ph", "sig", "tau"))
        with h5py.File(fname, 'w') as hf:
            hf.create_dataset('spin', data=self.a)
            hf.create_dataset('inc', data=self.th_o)
            hf.create_dataset('r_o', data=self.r_o)
            hf.create_dataset('alpha', data=self.alpha)
            hf.create_dataset('beta', data=self.beta)
            hf.create_dataset('t', data=self.t_s)
            hf.create_dataset('r', data=self.r_s)
            hf.create_dataset('theta', data=self.th_s)
            hf.create_dataset('phi', data=self.ph_s)
            hf.create_dataset('affine', data=self.sig_s)
            hf.create_dataset('mino', data=self.tausteps)
        return fname


def _require_h5py():
    try:
        import h5py
    except ImportError as exc:
        raise ImportError("savegeos / loadgeos need h5py, which is not installed (kgeo_b200 has no HDF5 writer of its own)") from exc
    return h5py


def loadgeos(infile):
    """Geodesics object from a file written by savegeos (kgeo/kerr_raytracing_utils.py:223-240); host (NumPy) arrays."""
    h5py = _require_h5py()
    with h5py.File(infile, 'r') as hf:
        a = hf['spin'][()]
        r_o = hf['r_o'][()]
        th_o = hf['inc'][()]
        alpha = hf['alpha'][()]
        beta = hf['beta'][()]
        mino_times = hf['mino'][()]
        affine_times = hf['affine'][()]
        coords = [hf[k][()] for k in ('t', 'r', 'theta', 'phi')]
    return Geodesics(a, [0, r_o, th_o, 0], [alpha, beta], mino_times, affine_times, coords)


def _lam_eta_arrays(lam, eta):
    if not isinstance(lam, np.ndarray):
        lam = np.array([lam]).flatten()
    if not isinstance(eta, np.ndarray):
        eta = np.array([eta]).flatten()
    if len(lam) != len(eta):
        raise Exception("lam, eta are different shapes!")
    return np.asarray(lam, dtype=np.float64), np.asarray(eta, dtype=np.float64)


def my_sign(x):
    """sign with sign(0) = +1 (kgeo/kerr_raytracing_utils.py:857-864)"""
    x = np.asarray(x)
    return np.where(x < 0, -1.0, 1.0)


def uplus_uminus(a, th_o, lam, eta):
    """(u_plus, u_minus, uratio, a2u_minus) -- kgeo/kerr_raytracing_utils.py:276-328"""
    _checks.check_spin(a)
    lam, eta = _lam_eta_arrays(lam, eta)
    out = _engine.ray_summary_lam_eta(a, 1000., th_o, lam, eta, want=("uplus", "uminus", "uratio", "a2uminus"))
    return tuple(_np(out[k]) for k in ("uplus", "uminus", "uratio", "a2uminus"))


def angular_turning(a, th_o, lam, eta):
    """(u_plus, u_minus, th_plus, th_minus, thclass) -- kgeo/kerr_raytracing_utils.py:242-274"""
    _checks.check_spin(a)
    lam, eta = _lam_eta_arrays(lam, eta)
    out = _engine.ray_summary_lam_eta(a, 1000., th_o, lam, eta, want=("uplus", "uminus", "thplus", "thminus"))
    thclass = np.zeros(lam.shape).astype(int)
    thclass[eta > 0] = 1
    thclass[eta < 0] = 2
    thclass[eta == 0] = 3
    return (_np(out["uplus"]), _np(out["uminus"]), _np(out["thplus"]), _np(out["thminus"]), thclass)


def radial_roots(a, lam, eta):
    """(r1, r2, r3, r4, rclass) with complex128 roots -- kgeo/kerr_raytracing_utils.py:330-380"""
    _checks.check_spin(a)
    lam, eta = _lam_eta_arrays(lam, eta)
    out = _engine.ray_summary_lam_eta(a, 1000., 0.25 * np.pi, lam, eta, want=("roots", "rclass"))
    roots = _np(out["roots"])
    r = [roots[k] + 1j * roots[4 + k] for k in range(4)]
    return (r[0], r[1], r[2], r[3], _np(out["rclass"]).astype(int))


def mino_total(a, r_o, eta, r1, r2, r3, r4):
    """Total Mino time to the horizon / back to infinity -- kgeo/kerr_raytracing_utils.py:382-548"""
    _checks.check_spin(a)
    if not (isinstance(r_o, float) and (r_o >= 100)):
        raise Exception("r_o should be a float > 100")
    eta = np.atleast_1d(np.asarray(eta, dtype=np.float64))
    roots = np.stack([np.real(np.atleast_1d(r)) for r in (r1, r2, r3, r4)] +
                     [np.imag(np.atleast_1d(r)) for r in (r1, r2, r3, r4)]).astype(np.float64)
    return _np(_engine.mino_total_from_roots(a, r_o, eta, roots))


def is_outside_crit(a, th_o, alpha, beta):
    """1 outside the critical curve, else 0 -- kgeo/kerr_raytracing_utils.py:794-826"""
    _checks.check_spin(a)
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    out = _engine.ray_summary(a, 1000., th_o, alpha, beta, want=("outside",))
    return _np(out["outside"])


def _counts(a, th_o, alpha, beta, tau, which):
    _checks.check_spin(a)
    alpha, beta = _checks.pixel_arrays(alpha, beta)
    shape = tuple(tau.shape)
    if shape[-1] != len(alpha):
        raise Exception("tau has incompatible shape in n_poloidal_orbits!" if which == "pol"
                        else "tau has incompatible shape in n_equatorial_crossings!")
    bundle = _engine.RayBundle(a, 1000., th_o, alpha, beta, clamp_lowspin=False, mode=1)
    neq, npol = bundle.crossing_counts(tau, want_eq=(which == "eq"), want_pol=(which == "pol"))
    res = neq if which == "eq" else npol
    return _np(res).reshape(shape)


def n_poloidal_orbits(a, th_o, alpha, beta, tau):
    """Fractional number of poloidal orbits at Mino time tau (GL19b Eq 35) -- kgeo/kerr_raytracing_utils.py:551-603"""
    return _counts(a, th_o, alpha, beta, tau, "pol")


def n_angular_turnings(a, th_o, alpha, beta, tau):
    """the fractional number of angular turnings at Mino time tau -- kgeo/kerr_raytracing_utils.py:686-792 (same checks and
    messages; the reference's own comments call its vortical branch unfinished: it is reproduced as it is).  Per-ray constants
    (u+-, F0, K) from the device library, the per-sample expression is linear in tau."""
    if not (isinstance(a, float) and (0 <= np.abs(a) < 1)):
        raise Exception("|a| should be a float in range [0,1)")
    if not (isinstance(th_o, float) and (0 <= th_o <= np.pi / 2.)):
        raise Exception("th_o should be a float in range (0,pi/2]")
    if not isinstance(alpha, np.ndarray):
        alpha = np.array([alpha]).flatten()
    if not isinstance(beta, np.ndarray):
        beta = np.array([beta]).flatten()
    if len(alpha) != len(beta):
        raise Exception("alpha, beta are different lengths!")
    alltau = len(tau.shape) != 1
    tau2 = tau if alltau else tau.reshape(1, tau.shape[0])
    if not (tau2.shape[1] == len(alpha)):
        raise Exception("tau has incompatible shape in n_poloidal_orbits!")
    ellipkinc = lambda phi, m: _engine.elementwise("kgeo_special_ellipkinc", (phi, m))
    lam = -alpha * np.sin(th_o)
    eta = (alpha ** 2 - a ** 2) * np.cos(th_o) ** 2 + beta ** 2
    out = np.empty(tau2.shape)
    vort = eta <= 0
    for vortical, mask in ((False, ~vort), (True, vort)):
        if not np.any(mask):
            continue
        up, um, ratio, a2um = uplus_uminus(a, th_o, lam[mask], eta[mask])
        t, b = tau2[:, mask], beta[mask]
        if vortical:
            m = 1 - ratio
            F0 = ellipkinc(np.arcsin(np.sqrt((np.cos(th_o) ** 2 - um) / (up - um))), m)
            w = np.sqrt(a2um)
        else:
            m = ratio
            F0 = ellipkinc(np.arcsin(np.cos(th_o) / np.sqrt(up)), m)
            w = np.sqrt(-a2um)
        K = ellipkinc(np.full(m.shape, 0.5 * np.pi), m)
        first = (t * w - F0) / (2 * K) + 0.5
        if vortical:
            first = np.where(first > 0.5, first + 0.5, first)
        out[:, mask] = np.where(b <= 0, first, (t * w + F0) / (2 * K) + 0.5)
    return out if alltau else out[0]


def n_equatorial_crossings(a, th_o, alpha, beta, tau):
    """Number of equatorial crossings up to Mino time tau (GL19b Eq 82) -- kgeo/kerr_raytracing_utils.py:605-684"""
    return _counts(a, th_o, alpha, beta, tau, "eq")
