"""Multi-GPU check, launched as  torchrun --nproc-per-node N tests/dist_check.py  on one box (NCCL):
row-sharded make_image gathered to rank 0 == single-GPU make_image bit for bit, and every rank's track shard == the same
rows of a whole-plane trace.  Prints 'dist_check ok' on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import kgeo_b200 as kg
    from kgeo_b200 import sharding
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    a, th_o = 0.94, 17 * np.pi / 180
    kw = dict(emissivity=kg.Emissivity('ring'), velocity=kg.Velocity('general', fac_subkep=0.7, beta_phi=0.7, beta_r=0.7),
              bfield=kg.Bfield('bz_guess', C=1.), polarization=True, specind=1.)
    # the gather fused into the image kernel (peer stores over NVLink), interleaved rows and ragged blocks of rows; then the
    # same through NCCL
    res = sharding.make_image_sharded(a, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 250, dst=0, **kw)
    res_blocks = sharding.make_image_sharded(a, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 250, dst=0, interleave=False, **kw)
    mode = "peer stores" if any(sharding._PEER_CACHE.values()) else "nccl (no peer mapping)"
    os.environ["KGEO_B200_GATHER"] = "nccl"
    res_nccl = sharding.make_image_sharded(a, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 250, dst=0, **kw)
    del os.environ["KGEO_B200_GATHER"]
    x = np.linspace(-8, 8, 120)
    A, B = [v.flatten() for v in np.meshgrid(x, x)]
    geos, (r0, r1) = sharding.raytrace_sharded(a, [0, 1000., th_o, 0], [A, B], 120, ngeo=50)
    whole = kg.raytrace_ana(a, [0, 1000., th_o, 0], [A, B], ngeo=50, plotdata=False)
    ok = all(np.array_equal(w[:, r0 * 120:r1 * 120], p, equal_nan=True) for w, p in zip(whole.geo_coords, geos.geo_coords))
    # BASELINE cfg 5 partition: whole images dealt round-robin to the ranks, gathered in sweep order
    cfgs = [(aa, tt * np.pi / 180) for aa in (0.0, 0.5, 0.99) for tt in (5.0, 45.0, 85.0)]
    sweep = sharding.make_image_sweep_sharded(cfgs, np.inf, 2, -10., 10., -10., 10., 20. / 64, dst=0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        single = kg.make_image(a, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 250, **kw)
        for other in (res, res_blocks, res_nccl):
            for got, want in zip(other, single):
                assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True), "sharded image differs"
        assert int(flag.item()) == 1, "a track shard differs from the whole-plane trace"
        for (aa, tt), got in zip(cfgs, sweep):
            want = kg.make_image(aa, np.inf, tt, 2, -10., 10., -10., 10., 20. / 64)
            assert all(np.array_equal(g_, w_, equal_nan=True) for g_, w_ in zip(got, want)), "sweep image differs"
        print("dist_check ok: world %d, image %s (gather: %s), track shards bitwise equal, %d-config sweep equal"
              % (world, res[0].shape, mode, len(cfgs)))
    sharding.release_peer_buffers()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
