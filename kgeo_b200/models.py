"""Pack Velocity / Bfield / Emissivity objects into the POD kgeo_models descriptor of the C ABI."""
import numpy as np

from . import _lib


def load_cache_from_file(filename, columns):
    """The reference's CSV reader of the 'fromfile' models (kgeo/velocities.py:626-657, kgeo/bfields.py:749-787): a '#'
    header line naming the columns, then comma-separated rows; missing columns are zero.  Returns (radii, c1, c2, c3)."""
    with open(filename, 'r') as f:
        header = f.readline().strip()
        if header[0] != '#':
            raise Exception("file %s does not have a header!" % filename)
        header = [x.strip() for x in header[1:].split(',')]
    data = np.loadtxt(filename, delimiter=',', skiprows=1)
    radii = data[:, header.index('r')]
    cols = [data[:, header.index(c)] if c in header else np.zeros_like(radii) for c in columns]
    return (radii,) + tuple(cols)


def device_table(rows):
    """4 x n float64 device table (row 0 = radii) for kgeo_models.*_table_dev"""
    from . import _engine
    return _engine.to_device(np.ascontiguousarray(np.vstack(rows), dtype=np.float64))


def pack_models(emissivity=None, velocity=None, bfield=None):
    if velocity is not None and getattr(velocity, 'veltype', None) == 'driftframe':
        # the drift frame is the E x B drift of ONE field model: kgeo_models carries a single bfield descriptor
        vb = velocity.bfield
        if bfield is None:
            bfield = vb
        elif bfield is not vb and (bfield.descriptor() != vb.descriptor() or bfield.table() is not vb.table()):
            raise NotImplementedError("Velocity('driftframe', bfield=X) is accelerated only together with bfield=X")
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
    # 'fromfile' models: device tables owned by the model objects (kept alive by them)
    for obj, prefix in ((velocity, "velocity"), (bfield, "bfield")):
        tab = obj.table() if (obj is not None and hasattr(obj, "table")) else None
        if tab is not None:
            setattr(M, prefix + "_table_dev", tab.data_ptr())
            setattr(M, prefix + "_table_n", int(tab.shape[1]))
    return M
