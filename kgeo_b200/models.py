"""Pack Velocity / Bfield / Emissivity objects into the POD kgeo_models descriptor of the C ABI."""
from . import _lib


def pack_models(emissivity=None, velocity=None, bfield=None):
    M = _lib.Models()
    M.velocity_type, M.bfield_type, M.emissivity_type = _lib.VEL['zamo'], _lib.BF['simple'], _lib.EM['constant']
    M.bfield_p[0] = 1.0
    M.emiscut_in, M.emiscut_out = 0.0, 1.e10
    for obj, prefix in ((velocity, "velocity"), (bfield, "bfield"), (emissivity, "emissivity")):
        if obj is None:
            continue
        tid, params = obj.descriptor()
        setattr(M, prefix + "_type", tid)
        arr = getattr(M, prefix + "_p")
        for i in range(8):
            arr[i] = params[i] if i < len(params) else 0.0
    if emissivity is not None:
        M.emiscut_in = float(emissivity.emiscut_in)
        M.emiscut_out = float(emissivity.emiscut_out)
    return M
