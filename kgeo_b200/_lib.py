"""ctypes binding of libkgeo_b200.so (the C ABI of include/kgeo_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the first compute symbol,
and every compute entry point returns a CUDA error when no device is present.
"""
import ctypes
import os

from . import build as _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

KGEO_E_BADARG, KGEO_E_RANGE, KGEO_E_MODEL = -1, -2, -3
MODE_ANY_ETA_POS, MODE_ANY_ETA_NONPOS, MODE_CLAMP_LOWSPIN_ETA, MODE_INPUT_LAM_ETA = 1, 2, 4, 8
MODE_POLISH_ROOTS, MODE_CARLSON_ONLY = 16, 32


class Observer(ctypes.Structure):
    _fields_ = [("a", ctypes.c_double), ("r_o", ctypes.c_double), ("th_o", ctypes.c_double)]


class Models(ctypes.Structure):
    _fields_ = [("velocity_type", ctypes.c_int32), ("bfield_type", ctypes.c_int32), ("emissivity_type", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("velocity_p", ctypes.c_double * 8), ("bfield_p", ctypes.c_double * 8),
                ("emissivity_p", ctypes.c_double * 8), ("emiscut_in", ctypes.c_double), ("emiscut_out", ctypes.c_double),
                ("velocity_table_dev", ctypes.c_void_p), ("bfield_table_dev", ctypes.c_void_p),
                ("velocity_table_n", ctypes.c_int64), ("bfield_table_n", ctypes.c_int64)]


class ImageOpts(ctypes.Structure):
    _fields_ = [("mbar_min", ctypes.c_int32), ("mbar_max", ctypes.c_int32), ("polarization", ctypes.c_int32),
                ("pathlength", ctypes.c_int32), ("specind", ctypes.c_double), ("diskangle", ctypes.c_double),
                ("nu_obs", ctypes.c_double)]


class ShardMap(ctypes.Structure):
    _fields_ = [("cols", ctypes.c_uint64), ("row_step", ctypes.c_uint64), ("row_offset", ctypes.c_uint64),
                ("plane_stride", ctypes.c_uint64), ("reserved", ctypes.c_uint64)]


class PeerHandle(ctypes.Structure):
    _fields_ = [("bytes", ctypes.c_ubyte * 64)]


VEL = dict(zamo=0, infall=1, kep=2, subkep=3, general=4, gelles=5, simfit=6, fromfile=7, driftframe=8)
BF = dict(simple=0, simple_rm1=1, bz_guess=2, gen_power=3, monopoleA=4, fromfile=5, bz_monopole=6, bz_para=7, power=8)
EM = dict(constant=0, bpl=1, glm=2, ring=3, thermal=4)

_P, _S, _V = c_double_p, ctypes.c_size_t, ctypes.c_void_p
_OBS, _MOD, _OPT = ctypes.POINTER(Observer), ctypes.POINTER(Models), ctypes.POINTER(ImageOpts)

# every symbol include/kgeo_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "kgeo_abi_version": (ctypes.c_int, []),
    "kgeo_last_error": (ctypes.c_char_p, []),
    "kgeo_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "kgeo_ray_workspace_slots": (ctypes.c_int, []),
    "kgeo_ray_workspace_doubles": (ctypes.c_size_t, [_S]),
    "kgeo_ray_slot": (ctypes.c_int, [_S, _P, ctypes.c_int, _P, _V]),
    "kgeo_launch_count": (ctypes.c_longlong, []),
    "kgeo_special_rf": (ctypes.c_int, [_S, _P, _P, _P, _P, _V]),
    "kgeo_special_rd": (ctypes.c_int, [_S, _P, _P, _P, _P, _V]),
    "kgeo_special_rj": (ctypes.c_int, [_S, _P, _P, _P, _P, _P, _V]),
    "kgeo_special_rc": (ctypes.c_int, [_S, _P, _P, _P, _V]),
    "kgeo_special_ellipk": (ctypes.c_int, [_S, _P, _P, _V]),
    "kgeo_special_ellipkinc": (ctypes.c_int, [_S, _P, _P, _P, _V]),
    "kgeo_special_ellipeinc": (ctypes.c_int, [_S, _P, _P, _P, _V]),
    "kgeo_special_ellippi": (ctypes.c_int, [_S, _P, _P, _P, _P, _V]),
    "kgeo_special_ellipj": (ctypes.c_int, [_S, _P, _P, _P, _P, _P, _P, _V]),
    "kgeo_ray_setup": (ctypes.c_int, [_OBS, _S, _P, _P, ctypes.c_int, _P, c_int32_p, _V]),
    "kgeo_ray_setup_conserved": (ctypes.c_int, [_OBS, _S, _P, _P, _P, ctypes.c_int, ctypes.c_int, _P, c_int32_p, _V]),
    "kgeo_track_integrals": (ctypes.c_int, [_OBS, _S, _P, c_int32_p, _S, _P, _S, ctypes.c_int] + [_P] * 7 + [_S, _V]),
    "kgeo_ray_summary": (ctypes.c_int, [_OBS, _S, _P, _P, ctypes.c_int, _P, c_int32_p, _P, _P, _P, _P, _P, _P, _P, _P,
                                        _V]),
    "kgeo_mino_total": (ctypes.c_int, [_OBS, _S, _P, _P, _V]),
    "kgeo_tracks": (ctypes.c_int, [_OBS, _S, _P, c_int32_p, _S, ctypes.c_double, _P, _S, ctypes.c_int,
                                   _P, _P, _P, _P, _P, _P, _S, _V]),
    "kgeo_raytrace": (ctypes.c_int, [_OBS, _S, _P, _P, ctypes.c_int, _S, ctypes.c_double, _P, c_int32_p,
                                     _P, _P, _P, _P, _P, _P, _S, _V]),
    "kgeo_raytrace_host": (ctypes.c_int, [_OBS, _S, _P, _P, ctypes.c_int, _S, ctypes.c_double,
                                          _P, _P, _P, _P, _P, _P, ctypes.c_int]),
    "kgeo_crossing_counts": (ctypes.c_int, [_OBS, _S, _P, _P, c_int32_p, _S, _P, _P, _P, _V]),
    "kgeo_r_equatorial": (ctypes.c_int, [_OBS, _S, _P, _P, ctypes.c_int, ctypes.c_int, _P, _P, _P, _P, _P, _V]),
    "kgeo_rho_of_req": (ctypes.c_int, [_OBS, _S, _P, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_int, _P, _P, _P, c_int32_p, _V]),
    "kgeo_rho_library": (ctypes.c_int, [_S, _P, _P, _P, ctypes.c_int, ctypes.c_int, _S, _P, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_int, _P, c_int32_p, _V]),
    "kgeo_critical_curve": (ctypes.c_int, [ctypes.c_double] * 6 + [_S, _P, _P, _V]),
    "kgeo_equatorial_image": (ctypes.c_int, [_OBS, _MOD, _OPT, _S, _P, _P] + [_P] * 12 + [_V]),
    "kgeo_equatorial_image_sharded": (ctypes.c_int, [_OBS, _MOD, _OPT, _S, _P, _P, ctypes.POINTER(ShardMap)] + [_P] * 12 + [_V]),
    "kgeo_shard_scatter": (ctypes.c_int, [_P, _S, _S, ctypes.POINTER(ShardMap), _P, _V]),
    "kgeo_peer_alloc": (ctypes.c_int, [_S, ctypes.POINTER(_V), ctypes.POINTER(PeerHandle)]),
    "kgeo_peer_open": (ctypes.c_int, [ctypes.POINTER(PeerHandle), ctypes.POINTER(_V)]),
    "kgeo_peer_close": (ctypes.c_int, [_V]),
    "kgeo_peer_free": (ctypes.c_int, [_V]),
    "kgeo_equatorial_image_host": (ctypes.c_int, [_OBS, _MOD, _OPT, _S, _P, _P] + [_P] * 12 + [ctypes.c_int]),
    "kgeo_velocity_u_lab": (ctypes.c_int, [_MOD, ctypes.c_double, _S, _P, _P, _P, _P, _P, _V]),
    "kgeo_bfield_lab": (ctypes.c_int, [_MOD, ctypes.c_double, _S, _P, _P, _P, _P, _P, _V]),
    "kgeo_bfield_omega": (ctypes.c_int, [_MOD, ctypes.c_double, _S, _P, _P, _V]),
    "kgeo_emissivity_jrest": (ctypes.c_int, [_MOD, ctypes.c_double, _S, _P, _P, _P, _P, ctypes.c_double,
                                             ctypes.c_double, _P, _V]),
    "kgeo_fieldline_guesses": (ctypes.c_int, [_OBS, _S, _S, _P, _P, _P, _S, _P, _P, _P, c_int32_p, ctypes.c_int, ctypes.c_double,
                                              ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int, c_int32_p, c_int32_p,
                                              _P, _P, _V]),
    "kgeo_fieldline_solve": (ctypes.c_int, [_OBS, _S, _S, _P, _P, c_int32_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, _P,
                                            ctypes.c_double, ctypes.c_int, _P, _P, _P, _P, _P, _P, c_int32_p, c_int32_p, _V]),
    "kgeo_jet_stokes": (ctypes.c_int, [_OBS, _S, _S, _P, _P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                       _P, ctypes.c_int64, ctypes.c_double, ctypes.c_int, _P, _P, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _P, _P, _P, _P, _P,
                                       c_int32_p, _V]),
    "kgeo_jet_sort": (ctypes.c_int, [_S, _S, ctypes.c_int, _P, _P, _P, _P, _P, _P, _P, _V]),
    "kgeo_trim_memory": (ctypes.c_int, [ctypes.c_int]),
    "kgeo_fp64_peak_probe": (ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.POINTER(ctypes.c_float), _V]),
}

_lib = None


class KgeoError(Exception):
    """Raised when a C-ABI call returns a non-zero status."""


def library_path():
    # KGEO_B200_LIB lets a developer point at an alternative build of the SAME library (kernel tuning experiments)
    return os.environ.get("KGEO_B200_LIB") or _build.LIBPATH


def lib():
    """Load (once) and return the shared library.  Never builds implicitly and never falls back."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError("kgeo_b200: %s is missing; build it with `python -m kgeo_b200.build` "
                              "(there is no CPU fallback)" % path)
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.kgeo_abi_version() != 4:
            raise ImportError("kgeo_b200: ABI version mismatch")
        _lib = L
    return _lib


def check(status):
    if status != 0:
        msg = lib().kgeo_last_error().decode()
        raise KgeoError("%s (status %d)" % (msg, status))


def launch_count():
    return int(lib().kgeo_launch_count())
