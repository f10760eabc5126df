"""Row sharding of the image plane over the GPUs of one box (one process per GPU, torch.distributed).

The per-pixel computation has no cross-pixel dependence, so there is NO data-path collective: every rank traces its own
rows.  The only communication is the final gather of the (small) image planes to one rank, and on the GPUs of a box that is
not a separate step either: the destination planes live in a PeerBuffer on the gathering rank's GPU, every other rank maps it
(CUDA IPC) and its image kernel stores each result straight to its place there over NVLink (kgeo_equatorial_image_sharded) --
no staging copy, no collective, no un-interleave pass; one barrier orders the reader after the writers.  Where peer mapping is
not available (gloo in the CPU tests, GPUs without peer access) the planes go through one all_gather_into_tensor
(gather_interleaved_rows) or a padded gather.  Geodesic tracks stay sharded on the devices that produced them
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


class PeerUnavailable(RuntimeError):
    """The GPUs of the group cannot map each other's memory (no peer access / no CUDA IPC in this container)."""


class _DeviceSpan(object):
    """`__cuda_array_interface__` view of raw device memory (float64), for torch.as_tensor."""

    def __init__(self, ptr, numel, owner):
        self.__cuda_array_interface__ = dict(shape=(int(numel),), typestr="<f8", data=(int(ptr), False), version=3, strides=None)
        self._owner = owner


class PeerBuffer(object):
    """`numel` float64 of device memory on rank `root`'s GPU that EVERY rank of the group can store to: the root allocates it
    (kgeo_peer_alloc), its 64-byte IPC handle is broadcast, the others map it into their own GPU's address space
    (kgeo_peer_open: NVLink peer access).  `ptr` is the address valid on the calling rank; `tensor()` is the root's view.
    Collective: every rank of the group constructs it, and closes it, together.  Raises PeerUnavailable on EVERY rank when any
    rank failed to map the buffer.  Without a process group it is simply local memory."""

    def __init__(self, numel, root=0, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _engine, _lib
        L = _lib.lib()
        _engine.require_cuda()
        self.numel, self.root, self.group = int(numel), int(root), group
        self.have_pg = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.have_pg else 0
        self.is_root = (self.rank == self.root) or not self.have_pg
        self.dev = _engine.device()
        self.ptr = None
        self._lib = L
        ok, err = 1, ""
        handle = _lib.PeerHandle()
        p = ctypes.c_void_p()
        with torch.cuda.device(self.dev):
            if self.is_root:
                _lib.check(L.kgeo_peer_alloc(self.numel * 8, ctypes.byref(p), ctypes.byref(handle)))
                self.ptr = int(p.value)
            if self.have_pg:
                box = [bytes(handle.bytes) if self.is_root else None]
                dist.broadcast_object_list(box, src=dist.get_global_rank(group, self.root) if group is not None else self.root,
                                           group=group)
                if not self.is_root:
                    handle = _lib.PeerHandle((ctypes.c_ubyte * 64).from_buffer_copy(box[0]))
                    if L.kgeo_peer_open(ctypes.byref(handle), ctypes.byref(p)) == 0:
                        self.ptr = int(p.value)
                    else:
                        ok, err = 0, L.kgeo_last_error().decode()
                flag = torch.tensor([ok], device=self.dev, dtype=torch.int32)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                if int(flag.item()) != 1:
                    self._release()
                    raise PeerUnavailable("PeerBuffer: a rank could not map the root's memory%s" % (": " + err if err else ""))

    def tensor(self, shape=None):
        """The root's float64 torch view of the buffer (None on the other ranks).  Valid until close(): the view does not own the
        memory."""
        if not self.is_root:
            return None
        import torch
        t = torch.as_tensor(_DeviceSpan(self.ptr, self.numel, self), device=self.dev)
        return t if shape is None else t.view(*shape)

    def barrier(self):
        """Order the ranks around a use of the buffer: called before the writers start (the root's previous read is over) and
        after their kernels were launched (NCCL's barrier runs behind them in stream order: the root then reads whole planes)."""
        if self.have_pg:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def _release(self):
        if self.ptr is not None:
            (self._lib.kgeo_peer_free if self.is_root else self._lib.kgeo_peer_close)(self.ptr)
            self.ptr = None

    def close(self):
        """Collective: the importers unmap, then the root frees."""
        import torch
        import torch.distributed as dist
        with torch.cuda.device(self.dev):
            torch.cuda.synchronize(self.dev)
            if not self.is_root:
                self._release()
            if self.have_pg:
                dist.barrier(group=self.group)
            self._release()


_PEER_CACHE = {}


def peer_buffer(numel, root=0, group=None):
    """A cached PeerBuffer of at least `numel` float64 for (root, group) (mapping costs milliseconds; a sweep of images reuses
    one), or None when the group cannot map peer memory (the result is remembered: the callers then gather through NCCL).
    Collective."""
    import os
    import torch.distributed as dist
    key = (int(root), id(group) if group is not None else None)
    have = _PEER_CACHE.get(key)
    if have is False or os.environ.get("KGEO_B200_GATHER", "peer") == "nccl":
        return None
    if have is not None and have.numel >= numel:
        return have
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) != "nccl":
        _PEER_CACHE[key] = False
        return None
    if have is not None:
        have.close()
    try:
        _PEER_CACHE[key] = PeerBuffer(numel, root, group)
    except PeerUnavailable:
        _PEER_CACHE[key] = False
        return None
    return _PEER_CACHE[key]


def release_peer_buffers():
    """Collective: close every cached PeerBuffer."""
    for key, buf in list(_PEER_CACHE.items()):
        if buf:
            buf.close()
        del _PEER_CACHE[key]


def to_host_pinned(tensor):
    """One asynchronous D2H copy of a CUDA tensor into page-locked memory (torch's caching pinned allocator), as a NumPy array."""
    import torch
    try:
        dst = torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
    except RuntimeError:
        dst = torch.empty(tensor.shape, dtype=tensor.dtype)
    dst.copy_(tensor, non_blocking=True)
    torch.cuda.current_stream(tensor.device).synchronize()
    return dst.numpy()


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


def _image_kwargs(kwargs):
    from . import equatorial_images as ei
    return (kwargs.get("emissivity", ei.emis_default), kwargs.get("velocity", ei.vel_default),
            kwargs.get("bfield", ei.bfield_default), kwargs.get("polarization", False), bool(kwargs.get("pathlength", False)),
            kwargs.get("specind", ei.SPECIND), kwargs.get("diskangle", ei.DISKANGLE), kwargs.get("nu_obs", ei.OBSFREQ))


def plane_layout(norder, npix, pathlength=False):
    """Layout of one image in a destination buffer: name -> element offset of its first plane, for the norder-plane outputs
    (I, Q, U, r, Ir, g, sinthb[, lp, Ie]) followed by the single planes nmax, npol; and the total number of elements."""
    names = ["I", "Q", "U", "r", "Ir", "g", "sinthb"] + (["lp", "Ie"] if pathlength else [])
    off = {k: i * norder * npix for i, k in enumerate(names)}
    off["nmax"] = len(names) * norder * npix
    off["npol"] = off["nmax"] + npix
    return off, off["npol"] + npix


def _image_tuple(flat, off, norder, npix, pathlength):
    """make_image's return tuple from one image of a host buffer laid out by plane_layout (the reference's (npix, mbar_max + 1)
    indexing of the order planes, kgeo/equatorial_images.py:113-121)."""
    plane = lambda k: flat[off[k]:off[k] + norder * npix].reshape(norder, npix).T
    res = tuple(plane(k) for k in ("I", "Q", "U", "r", "Ir", "g", "sinthb")) + (flat[off["nmax"]:off["nmax"] + npix],
                                                                               flat[off["npol"]:off["npol"] + npix])
    return res + (plane("lp"), plane("Ie")) if pathlength else res


def make_image_sharded(a, r_o, th_o, mbar_max, alpha_min, alpha_max, beta_min, beta_max, psize, dst=0, group=None,
                       interleave=True, **kwargs):
    """make_image with the pixel rows sharded over the ranks of `group`; rank `dst` returns the reference-shaped tuple
    (as equatorial_images.make_image), other ranks return None.  Needs an initialised process group.  On the GPUs of a box the
    gather is fused into the image kernel (peer stores into rank dst's planes, one barrier); otherwise NCCL / gloo gathers."""
    import torch
    import torch.distributed as dist
    from . import equatorial_images as ei
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    alpha_arr, beta_arr = ei.pixel_grid(alpha_min, alpha_max, beta_min, beta_max, psize)
    n_alpha = int(np.floor(alpha_max - alpha_min) / psize)
    nrows = alpha_arr.size // n_alpha
    if interleave:
        al, be, rows = shard_pixels_interleaved(alpha_arr, beta_arr, n_alpha, rank, world)
        step, offset = world, rank
    else:
        al, be, rows = shard_pixels(alpha_arr, beta_arr, n_alpha, rank, world)
        step, offset = 1, rows[0]
    model = _image_kwargs(kwargs)
    pathlength = model[4]
    norder, npix = mbar_max + 1, nrows * n_alpha
    off, total = plane_layout(norder, npix, pathlength)
    peer = peer_buffer(total, dst, group)
    if peer is not None:
        peer.barrier()
        if al.size:
            ei.image_into(a, r_o, th_o, 0, mbar_max, al, be, *model, dst={k: peer.ptr + 8 * o for k, o in off.items()},
                          cols=n_alpha, row_step=step, row_offset=offset, plane_stride=npix)
        peer.barrier()
        if rank != dst:
            return None
        return _image_tuple(to_host_pinned(peer.tensor()[:total]), off, norder, npix, pathlength)
    out = ei.image_device(a, r_o, th_o, 0, mbar_max, al, be, *model)
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
    (single GPU: every configuration is local).  On the GPUs of a box every image is stored by its kernel straight into rank
    dst's [nconf, planes] buffer (peer stores), which then goes to page-locked host memory in one copy."""
    import torch
    import torch.distributed as dist
    from . import equatorial_images as ei
    from . import _engine
    have_pg = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if have_pg else 1
    rank = dist.get_rank(group) if have_pg else 0
    alpha_arr, beta_arr = ei.pixel_grid(alpha_min, alpha_max, beta_min, beta_max, psize)
    model = _image_kwargs(kwargs)
    pathlength = model[4]
    dev = _engine.device()
    npix, norder, nconf = alpha_arr.size, mbar_max + 1, len(configs)
    off, per = plane_layout(norder, npix, pathlength)
    peer = peer_buffer(nconf * per, dst, group) if (have_pg and world > 1) else None
    if peer is not None or world == 1:
        al, be = _engine.to_device(alpha_arr, dev), _engine.to_device(beta_arr, dev)
        if peer is None:
            local = _engine.empty((nconf * per,), dev)
            base = local.data_ptr()
        else:
            peer.barrier()
            base = peer.ptr
        for i in config_indices(nconf, rank, world):
            a, th_o = configs[i]
            ei.image_into(float(a), r_o, float(th_o), 0, mbar_max, al, be, *model,
                          dst={k: base + 8 * (i * per + o) for k, o in off.items()}, cols=npix, row_step=1, row_offset=0,
                          plane_stride=npix)
        if peer is not None:
            peer.barrier()
            if rank != dst:
                return None
            local = peer.tensor()[:nconf * per]
        flat = to_host_pinned(local)
        return [_image_tuple(flat[c * per:(c + 1) * per], off, norder, npix, pathlength) for c in range(nconf)]
    names = ["I", "Q", "U", "r", "Ir", "g", "sinthb"] + (["lp", "Ie"] if pathlength else [])
    planes, scal = [], []
    for i in config_indices(nconf, rank, world):
        a, th_o = configs[i]
        out = ei.image_device(float(a), r_o, float(th_o), 0, mbar_max, alpha_arr, beta_arr, *model)
        planes.append(torch.stack([out[k] for k in names]))              # [nplane, norder, npix]
        scal.append(torch.stack([out["nmax"], out["npol"]]))             # [2, npix]
    P = torch.stack(planes) if planes else torch.zeros((0, len(names), norder, npix), dtype=torch.float64, device=dev)
    S = torch.stack(scal) if scal else torch.zeros((0, 2, npix), dtype=torch.float64, device=dev)
    P = gather_configs(P, nconf, dst, group)
    S = gather_configs(S, nconf, dst, group)
    if rank != dst:
        return None
    P, S = to_host_pinned(P), to_host_pinned(S)
    res = []
    for c in range(nconf):
        tup = tuple(P[c, i].T for i in range(7)) + (S[c, 0], S[c, 1])
        if pathlength:
            tup = tup + (P[c, 7].T, P[c, 8].T)
        res.append(tup)
    return res
