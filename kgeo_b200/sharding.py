"""Row sharding of the image plane over the GPUs of one box (one process per GPU, torch.distributed).

The per-pixel computation has no cross-pixel dependence, so there is NO data-path collective: every rank traces its own
contiguous block of image rows.  The only communication is the final gather of the (small) image planes to one rank
(one NCCL all_gather_into_tensor over NVLink on GPUs -- gather_interleaved_rows -- gloo in the CPU tests); geodesic tracks stay sharded on the devices that produced them
(1024^2 x 500 samples x 4 coordinates is 16.8 GB, 4096^2 is 268 GB -- they are consumed where they are).
"""
import numpy as np


def row_range(nrows, rank, world):
    """Contiguous block [r0, r1) of `nrows` image rows owned by `rank` of `world` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of size %d" % (rank, world))
    base, extra = divmod(int(nrows), int(world))
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def all_row_ranges(nrows, world):
    return [row_range(nrows, r, world) for r in range(world)]


def row_indices(nrows, rank, world, interleave=False):
    """Image rows owned by `rank`: a contiguous block (default) or every world-th row (`interleave=True`, better load
    balance: rows through the shadow are more expensive than rows of escaping rays)."""
    if interleave:
        if not (0 <= rank < world):
            raise ValueError("rank %d outside world of size %d" % (rank, world))
        return np.arange(rank, nrows, world)
    r0, r1 = row_range(nrows, rank, world)
    return np.arange(r0, r1)


def shard_pixels_interleaved(alpha, beta, ncols, rank, world):
    """As shard_pixels, for the interleaved distribution; returns (alpha_local, beta_local, row_index_array)."""
    alpha = np.asarray(alpha).reshape(-1, ncols)
    beta = np.asarray(beta).reshape(-1, ncols)
    ix = row_indices(alpha.shape[0], rank, world, interleave=True)
    return np.ascontiguousarray(alpha[ix].ravel()), np.ascontiguousarray(beta[ix].ravel()), ix


def shard_pixels(alpha, beta, ncols, rank, world):
    """Slice flattened row-major pixel arrays (row = constant beta, as produced by np.meshgrid(alphas, betas)) to the
    rows owned by `rank`.  Returns (alpha_local, beta_local, (r0, r1))."""
    alpha = np.asarray(alpha)
    beta = np.asarray(beta)
    if alpha.shape != beta.shape or alpha.ndim != 1 or alpha.size % ncols:
        raise ValueError("alpha/beta must be flattened [nrows*ncols] arrays")
    nrows = alpha.size // ncols
    r0, r1 = row_range(nrows, rank, world)
    return alpha[r0 * ncols:r1 * ncols], beta[r0 * ncols:r1 * ncols], (r0, r1)


def gather_interleaved_rows(local, gathered, final, rows_local, ncols, group=None):
    """The final gather of a row-interleaved image (row % world == rank), without padding, Python loops or temporaries:
    ONE all_gather_into_tensor of every rank's plane-major shard `local` [nplane, rows_local * ncols] into the preallocated
    `gathered` [world, nplane, rows_local * ncols] (NCCL over NVLink), then -- on ranks that pass `final`
    [nplane, rows_local * world * ncols] -- one strided device copy that puts local row lr of rank r at image row
    lr * world + r.  Needs equal shards (nrows % world == 0); gather_rows handles the ragged case."""
    import torch.distributed as dist
    world = gathered.shape[0]
    nplane = local.shape[0]
    dist.all_gather_into_tensor(gathered.view(-1), local.reshape(-1), group=group)
    if final is not None:
        final.view(nplane, rows_local, world, ncols).copy_(gathered.view(world, nplane, rows_local, ncols).permute(1, 2, 0, 3))
    return final


def gather_rows(local, nrows, ncols, dst=0, group=None, interleave=False):
    """Gather row-sharded planes to rank `dst` (rows distributed as by row_indices).

    local: tensor [..., rows_local*ncols] (any leading plane dimensions) on every rank, laid out row-major in the pixel
    index.  Returns the assembled [..., nrows*ncols] tensor on `dst` and None elsewhere.  Uses all_gather on equal-size
    padded buffers when the split is uneven (NCCL has no ragged gather) -- the planes are small (184 B/pixel).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = [row_indices(nrows, r, world, interleave) for r in range(world)]
    lead = tuple(local.shape[:-1])
    if interleave and nrows % world == 0 and local.is_contiguous():     # equal shards: one collective, one strided copy
        rl = nrows // world
        flat = local.reshape(-1, rl * ncols)
        gathered = torch.empty((world,) + tuple(flat.shape), dtype=local.dtype, device=local.device)
        final = torch.empty((flat.shape[0], nrows * ncols), dtype=local.dtype, device=local.device) if rank == dst else None
        gather_interleaved_rows(flat, gathered, final, rl, ncols, group)
        return final.reshape(lead + (nrows * ncols,)) if rank == dst else None
    maxrows = max(len(ix) for ix in rows)
    buf = torch.zeros(lead + (maxrows * ncols,), dtype=local.dtype, device=local.device)
    buf[..., :local.shape[-1]] = local
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
    else:
        parts = None
    if dist.get_backend(group) == "nccl":
        # NCCL gather: all ranks send to dst over NVLink
        dist.gather(buf, gather_list=parts, dst=dst, group=group)
    else:
        dist.gather(buf, gather_list=parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty(lead + (nrows, ncols), dtype=local.dtype, device=local.device)
    for ix, p in zip(rows, parts):
        if len(ix):
            out[..., torch.as_tensor(ix, device=local.device), :] = p[..., :len(ix) * ncols].reshape(lead + (len(ix), ncols))
    return out.reshape(lead + (nrows * ncols,))


def make_image_sharded(a, r_o, th_o, mbar_max, alpha_min, alpha_max, beta_min, beta_max, psize, dst=0, group=None,
                       interleave=True, **kwargs):
    """make_image with the pixel rows sharded over the ranks of `group`; rank `dst` returns the reference-shaped tuple
    (as equatorial_images.make_image), other ranks return None.  Needs an initialised NCCL process group."""
    import torch
    import torch.distributed as dist
    from . import equatorial_images as ei
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    alpha_arr, beta_arr = ei.pixel_grid(alpha_min, alpha_max, beta_min, beta_max, psize)
    n_alpha = int(np.floor(alpha_max - alpha_min) / psize)
    nrows = alpha_arr.size // n_alpha
    if interleave:
        al, be, _ = shard_pixels_interleaved(alpha_arr, beta_arr, n_alpha, rank, world)
    else:
        al, be, _ = shard_pixels(alpha_arr, beta_arr, n_alpha, rank, world)
    pathlength = bool(kwargs.get("pathlength", False))
    out = ei.image_device(a, r_o, th_o, 0, mbar_max, al, be,
                          kwargs.get("emissivity", ei.emis_default), kwargs.get("velocity", ei.vel_default),
                          kwargs.get("bfield", ei.bfield_default), kwargs.get("polarization", False), pathlength,
                          kwargs.get("specind", ei.SPECIND), kwargs.get("diskangle", ei.DISKANGLE),
                          kwargs.get("nu_obs", ei.OBSFREQ))
    names = ["I", "Q", "U", "r", "Ir", "g", "sinthb"] + (["lp", "Ie"] if pathlength else [])
    planes = torch.stack([out[k] for k in names])                       # [nplane, norder, npix_local]
    scal = torch.stack([out["nmax"], out["npol"]])                      # [2, npix_local]
    planes = gather_rows(planes, nrows, n_alpha, dst, group, interleave)
    scal = gather_rows(scal, nrows, n_alpha, dst, group, interleave)
    if rank != dst:
        return None
    P = planes.cpu().numpy()
    S = scal.cpu().numpy()
    res = tuple(P[i].T for i in range(7)) + (S[0], S[1])
    if pathlength:
        res = res + (P[7].T, P[8].T)
    return res


def raytrace_sharded(a, observer_coords, image_coords, ncols, ngeo, maxtaufrac=1. - 1.e-10, group=None):
    """raytrace_ana on this rank's block of image rows; the returned Geodesics holds the LOCAL shard on this rank's GPU
    (tracks are never gathered).  Returns (geodesics, (row0, row1))."""
    import torch.distributed as dist
    from .kerr_raytracing_ana import raytrace_ana
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    al, be, rows = shard_pixels(image_coords[0], image_coords[1], ncols, rank, world)
    geos = raytrace_ana(a, observer_coords, [np.ascontiguousarray(al), np.ascontiguousarray(be)], ngeo=ngeo,
                        maxtaufrac=maxtaufrac, plotdata=False)
    return geos, rows


# ---- BASELINE cfg 5: a sweep of whole images, partitioned by CONFIG ------------------------------------
def config_indices(nconf, rank, world):
    """Configurations owned by `rank`: every world-th one (cost grows with spin and inclination along the usual sweep
    order, so a round-robin split balances better than blocks)."""
    return list(range(rank, nconf, world))


def gather_configs(local, nconf, dst=0, group=None):
    """Gather per-configuration tensors to rank `dst`.

    local: tensor [nconf_local, ...] holding this rank's configurations in the order of config_indices.  Returns the
    [nconf, ...] tensor in sweep order on `dst`, None elsewhere (equal-size padded buffers, one gather)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    owned = [config_indices(nconf, r, world) for r in range(world)]
    maxn = max(len(ix) for ix in owned)
    tail = tuple(local.shape[1:])
    buf = torch.zeros((maxn,) + tail, dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gather_list=parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty((nconf,) + tail, dtype=local.dtype, device=local.device)
    for ix, p in zip(owned, parts):
        if len(ix):
            out[torch.as_tensor(ix, device=local.device)] = p[:len(ix)]
    return out


def make_image_sweep_sharded(configs, r_o, mbar_max, alpha_min, alpha_max, beta_min, beta_max, psize, dst=0, group=None,
                             **kwargs):
    """Images of a list of (a, th_o) configurations (BASELINE cfg 5: 64 spin x inclination configs of 512 x 512 pixels over
    8 GPUs), one whole image per fused kernel launch, configurations dealt round-robin to the ranks of `group`; rank `dst`
    returns a list of make_image-shaped tuples in the order of `configs`, other ranks None.  Works without a process group
    (single GPU: every configuration is local)."""
    import torch
    import torch.distributed as dist
    from . import equatorial_images as ei
    have_pg = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if have_pg else 1
    rank = dist.get_rank(group) if have_pg else 0
    alpha_arr, beta_arr = ei.pixel_grid(alpha_min, alpha_max, beta_min, beta_max, psize)
    pathlength = bool(kwargs.get("pathlength", False))
    names = ["I", "Q", "U", "r", "Ir", "g", "sinthb"] + (["lp", "Ie"] if pathlength else [])
    planes, scal = [], []
    for i in config_indices(len(configs), rank, world):
        a, th_o = configs[i]
        out = ei.image_device(float(a), r_o, float(th_o), 0, mbar_max, alpha_arr, beta_arr,
                              kwargs.get("emissivity", ei.emis_default), kwargs.get("velocity", ei.vel_default),
                              kwargs.get("bfield", ei.bfield_default), kwargs.get("polarization", False), pathlength,
                              kwargs.get("specind", ei.SPECIND), kwargs.get("diskangle", ei.DISKANGLE),
                              kwargs.get("nu_obs", ei.OBSFREQ))
        planes.append(torch.stack([out[k] for k in names]))              # [nplane, norder, npix]
        scal.append(torch.stack([out["nmax"], out["npol"]]))             # [2, npix]
    from . import _engine
    dev = _engine.device()
    npix, norder = alpha_arr.size, mbar_max + 1
    P = torch.stack(planes) if planes else torch.zeros((0, len(names), norder, npix), dtype=torch.float64, device=dev)
    S = torch.stack(scal) if scal else torch.zeros((0, 2, npix), dtype=torch.float64, device=dev)
    if have_pg and world > 1:
        P = gather_configs(P, len(configs), dst, group)
        S = gather_configs(S, len(configs), dst, group)
        if rank != dst:
            return None
    P, S = P.cpu().numpy(), S.cpu().numpy()
    res = []
    for c in range(len(configs)):
        tup = tuple(P[c, i].T for i in range(7)) + (S[c, 0], S[c, 1])
        if pathlength:
            tup = tup + (P[c, 7].T, P[c, 8].T)
        res.append(tup)
    return res
