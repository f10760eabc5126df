"""CPU test of the N>1 host path: two gloo ranks shard an image plane by rows, 'compute' their rows, and gather the
planes to rank 0 with kgeo_b200.sharding.gather_rows (the same code path NCCL takes on GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nrows, ncols, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from kgeo_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, B = np.meshgrid(np.linspace(-1, 1, ncols), np.linspace(-2, 2, nrows))
        al, be, (r0, r1) = sharding.shard_pixels(A.flatten(), B.flatten(), ncols, rank, world)
        # stand-in for the per-pixel kernel: any pure function of the pixel coordinates
        planes = torch.from_numpy(np.stack([np.stack([al + 10 * be, al * be]), np.stack([al - be, al + be])]))  # [2,2,n]
        out = sharding.gather_rows(planes, nrows, ncols, dst=0)
        vec = sharding.gather_rows(torch.from_numpy(np.ascontiguousarray(al)), nrows, ncols, dst=0)
        # interleaved distribution (row % world == rank)
        al2, be2, ix = sharding.shard_pixels_interleaved(A.flatten(), B.flatten(), ncols, rank, world)
        assert np.array_equal(ix % world, np.full(len(ix), rank))
        out2 = sharding.gather_rows(torch.from_numpy(np.stack([al2 + 10 * be2, al2 * be2])), nrows, ncols, dst=0, interleave=True)
        if rank == 0:
            assert np.array_equal(out2.numpy(), out.numpy()[0])
            q.put((out.numpy(), vec.numpy()))
        else:
            assert out is None and vec is None and out2 is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nrows,ncols", [(2, 8, 5), (2, 7, 3), (3, 10, 4)])
def test_gather_rows_gloo(world, nrows, ncols):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nrows, ncols, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, vec = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    A, B = np.meshgrid(np.linspace(-1, 1, ncols), np.linspace(-2, 2, nrows))
    A, B = A.flatten(), B.flatten()
    want = np.stack([np.stack([A + 10 * B, A * B]), np.stack([A - B, A + B])])
    assert np.array_equal(out, want)
    assert np.array_equal(vec, A)


def _worker_cfg(rank, world, port, nconf, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from kgeo_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.config_indices(nconf, rank, world)
        # stand-in for one fused image per configuration: planes that encode the configuration index
        local = torch.tensor([[[10.0 * c + p + 0.25 * k for k in range(3)] for p in range(2)] for c in mine],
                             dtype=torch.float64).reshape(len(mine), 2, 3)
        out = sharding.gather_configs(local, nconf, dst=0)
        if rank == 0:
            q.put(out.numpy())
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nconf", [(2, 8), (2, 5), (3, 7)])
def test_gather_configs_gloo(world, nconf):
    """BASELINE cfg 5 partition (whole images dealt round-robin to ranks) and its gather, on gloo."""
    import torch.multiprocessing as mp
    from kgeo_b200 import sharding
    assert sorted(sum((sharding.config_indices(nconf, r, world) for r in range(world)), [])) == list(range(nconf))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_cfg, args=(r, world, port, nconf, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.array([[[10.0 * c + p + 0.25 * k for k in range(3)] for p in range(2)] for c in range(nconf)])
    assert np.array_equal(out, want)


@pytest.mark.parametrize("pathlength", [False, True])
def test_plane_layout_and_row_map_on_the_host(pathlength):
    """The destination layout of the fused image + gather (sharding.plane_layout / kgeo_shard_map): writing every shard's
    pixels at (lr * row_step + row_offset) * cols + c with NumPy assembles exactly the whole planes, for interleaved rows and
    for ragged blocks of rows; _image_tuple returns make_image's shapes."""
    sys.path.insert(0, ROOT)
    from kgeo_b200 import sharding
    nrows, ncols, norder, world = 11, 7, 3, 4
    npix = nrows * ncols
    off, total = sharding.plane_layout(norder, npix, pathlength)
    names = [k for k in off]
    assert total == (len(names) - 2) * norder * npix + 2 * npix
    rng = np.random.default_rng(5)
    whole = rng.standard_normal(total)
    for interleave in (True, False):
        dst = np.full(total, np.nan)
        for rank in range(world):
            rows = sharding.row_indices(nrows, rank, world, interleave)
            step, offset = (world, rank) if interleave else (1, rows[0] if len(rows) else 0)
            i = np.arange(len(rows) * ncols)
            d = (i // ncols * step + offset) * ncols + i % ncols          # the kernel's row map
            assert np.array_equal(d, (rows[:, None] * ncols + np.arange(ncols)).ravel())
            for k in names:
                for m in range(1 if k in ("nmax", "npol") else norder):
                    dst[off[k] + m * npix + d] = whole[off[k] + m * npix + d]
        assert np.array_equal(dst, whole)
    tup = sharding._image_tuple(whole, off, norder, npix, pathlength)
    assert len(tup) == (11 if pathlength else 9)
    assert tup[0].shape == (npix, norder) and tup[7].shape == (npix,)
    assert np.array_equal(tup[3][:, 2], whole[off["r"] + 2 * npix:off["r"] + 3 * npix])
