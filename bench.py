#!/usr/bin/env python
"""bench.py -- fp64 geodesic points/s of the analytic Kerr ray tracer (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

One "step" = one pass of the hot path over the workload.
  N = 1   BASELINE cfg 3: per-ray setup + track pass producing r, theta, phi, t for 1024 x 1024 pixels x 500 Mino-time samples
          (a = 0.94, theta_o = 17 deg, r_o = 1000, +-10 M).
  N > 1   BASELINE cfg 4, the north_star's target run: 4096 x 4096 pixels x 500 samples, image rows INTERLEAVED over the N ranks
          (row % N == rank), and per step on every rank: setup + track pass of its rows (r, theta, phi, t stay resident:
          268 GB / N per GPU), the fused n = 0..2 equatorial image of its rows, and the gather of the image planes to rank 0
          (one NCCL all_gather_into_tensor over NVLink + the row un-interleave on rank 0).  `value` counts the geodesic
          points of all ranks; image_ms and collective_ms are reported separately.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      FP64-pipe roofline of the track pass (hot kernel + fix-up kernel): achieved = executed FP64 flops per point (ncu
                DFMA/DMUL/DADD thread-instruction counts of the capture recorded in profiles/r02_ncu_summary.json -- used only
                while the kernel sources still hash to what that capture profiled, else null) x points / CUDA-event duration;
                peak = DFMA-chain rate measured live (MEASURED_PEAKS.json has no FP64 entry).  "canonical" restates the launch
                in SURVEY 8(d)'s work model of the reference's algorithm (2400 flop/point, Carlson duplication per sample).
  parity        (N = 1) the GPU output of the benched launch against the CPU oracle at 4096 cfg-3 pixels x 500 samples: per
                coordinate max / 99.9th percentile / fraction above a plain 1e-10, and the same with the stated gates
                (oracle/gates.py); the run FAILS if a gate breaks.  (N > 1) shard_check: sharded == whole, bit for bit.
  cpu_baseline  the CPU oracle port of the reference (oracle/kgeo_oracle.py, SciPy path) on this box's host cores.
  e2e           same metric through the host-buffer C-ABI call kgeo_raytrace_host (host alpha/beta in, r/theta/phi/t into
                pinned host arrays inside the timed region); e2e.python = the same through the Python drop-in
                raytrace_ana(...).geo_coords; e2e.ceiling_gbs = bare pinned D2H copies of the same size on all ranks at once.
  secondary     the fused n = 0..2 image kernel at 4096^2 with its own roofline / e2e / cpu_baseline; at N = 8 the cfg-5 sweep.
"""
import argparse
import contextlib
import ctypes
import hashlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPIN, INC_DEG, R_O, HALF_FOV = 0.94, 17.0, 1000.0, 10.0
NGEO = 500
MAXTAUFRAC = 1.0 - 1.0e-10
CANONICAL_FLOP_PER_POINT = 2400.0   # SURVEY.md 8(d) work model: Carlson duplication, the reference's algorithm
CANONICAL_FLOP_PER_PIXEL = 3000.0   # n = 0..2 image
BYTES_PER_POINT = 32.0              # r, theta, phi, t written once
IMG_PLANES = 24                     # I,Q,U,r,Ir,g,sinthb x 3 orders + nmax, npol, imax: doubles per pixel of an n=0..2 image
GATHER_DEFAULT = "copy_engine"     # N > 1: how the image planes reach rank 0 (see run_ours)
PARITY_PIXELS = 4096


def grid_shape(world):
    """(rows, cols) of the benched image plane: cfg 3 on one GPU, cfg 4 on several."""
    return (1024, 1024) if world == 1 else (4096, 4096)


def plane_pixels(world):
    rows, cols = grid_shape(world)
    x = np.linspace(-HALF_FOV, HALF_FOV, cols)
    y = np.linspace(-HALF_FOV, HALF_FOV, rows)
    return x, y


def workload_pixels(rank, world):
    """Rows rank, rank + world, ... of the plane spanning +-10 M.  Rows are interleaved so that every rank sees the
    same mix of root classes: rays through the shadow cost more than rays that escape."""
    x, y = plane_pixels(world)
    A, B = np.meshgrid(x, y[rank::world])
    return np.ascontiguousarray(A.flatten()), np.ascontiguousarray(B.flatten())


def workload_name(world):
    rows, cols = grid_shape(world)
    if world == 1:
        return ("BASELINE cfg 3 geodesic tracks: a=0.94, theta_o=17deg, r_o=1000, +-10M, %dx%d px x %d Mino-time samples -> "
                "r,theta,phi,t" % (rows, cols, NGEO))
    return ("BASELINE cfg 4: a=0.94, theta_o=17deg, +-10M, %dx%d px x %d Mino-time samples -> r,theta,phi,t (r_o=1000) + fused "
            "n=0..2 equatorial image (r_o=inf, bpl/zamo/rad) of the same pixels, rows interleaved over %d GPUs, image planes "
            "gathered to rank 0" % (rows, cols, NGEO, world))


def kernel_source_hash():
    """sha256 over the CUDA sources: ties the ncu-derived constants of profiles/r02_ncu_summary.json to the code they profiled."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "kgeo_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def ncu_summary():
    """profiles/r02_ncu_summary.json if it describes the kernels as they are now, else None."""
    try:
        s = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_summary.json")))
    except Exception:
        return None
    return s if s.get("csrc_sha256_16") == kernel_source_hash() else None


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's SciPy path (test infrastructure, used here only as the thing compared to)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    a, r_o, th_o, A, B, ngeo, keep = args
    np.seterr(all="ignore")
    from oracle import kgeo_oracle as ko
    tau, sig, (t, r, th, ph) = ko.raytrace(a, [0, r_o, th_o, 0], [A, B], ngeo, MAXTAUFRAC)
    if keep:
        return dict(tau=tau, t=t, r=r, th=th, ph=ph)
    return float(np.nansum(r[1]))


def sample_indices(npix, n):
    """n distinct pixels spread over the whole plane (257 is coprime to the power-of-two pixel counts: every column and
    row of the image is visited, unlike a power-of-two stride, which would walk down four columns only)."""
    return (np.arange(n, dtype=np.int64) * 257) % npix


def cpu_sample(npix_per_worker, workers, world=1, keep=False):
    x, y = plane_pixels(world)
    A, B = [v.flatten() for v in np.meshgrid(x, y)]
    idx = sample_indices(A.size, npix_per_worker * workers)
    th_o = INC_DEG * np.pi / 180
    return idx, A, B, [(SPIN, R_O, th_o, A[idx[w::workers]], B[idx[w::workers]], NGEO, keep) for w in range(workers)]


def time_cpu(npix_per_worker, workers, steps, warmup, world=1, keep=False):
    idx, A, B, chunks = cpu_sample(npix_per_worker, workers, world, keep)
    npts = sum(c[3].size for c in chunks) * NGEO
    if workers == 1:
        run = lambda: [_cpu_chunk(c) for c in chunks]
        pool = None
    else:
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(workers)
        run = lambda: pool.map(_cpu_chunk, chunks)
    try:
        for _ in range(warmup):
            run()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = run()
        dt = (time.perf_counter() - t0) / steps
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    return npts / dt, dt, npts, idx, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_worker = 512           # 512 px x 500 samples = 2.6e5 points per worker per step (~2-3 s of SciPy)
    value, dt, npts, _, _ = time_cpu(per_worker, workers, args.steps, args.warmup, world)
    rows, cols = grid_shape(world)
    sample = "%d pixels (spread over the %dx%d grid) x %d samples per step, %d processes" % (
        per_worker * workers, rows, cols, NGEO, workers)
    line = {
        "impl": "reference", "metric": "fp64 geodesic points/sec (raytrace_ana)", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(world),
                   "note": "reference algorithm (oracle port of kgeo's SciPy path; the Python reference itself cannot "
                           "travel to the GPU box) on the host cores, geodesic tracks of a bounded pixel sample"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler, affinity
# ---------------------------------------------------------------------------------------------------------------------
def local_cpu_affinity(gpu_index):
    """Bind this process to the CPUs nvidia-smi reports as local to the GPU, so that the pinned result buffers of the e2e
    leg are first-touched on the NUMA node next to the GPU's PCIe root.  Returns the CPU list description or None."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()
        header = [h.strip() for h in out[0].split("\t")]
        col = next(i for i, h in enumerate(header) if "CPU Affinity" in h)
        row = next(l for l in out[1:] if l.startswith("GPU%d" % gpu_index))
        spec = [c.strip() for c in row.split("\t")][col]
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import kgeo_b200                                   # fails loudly if libkgeo_b200.so is missing
    from kgeo_b200 import _engine, _lib, sharding
    from kgeo_b200 import equatorial_images as ei
    from kgeo_b200.models import pack_models

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    th_o = INC_DEG * np.pi / 180
    profile_only = bool(os.environ.get("KGEO_BENCH_PROFILE"))
    rows, cols = grid_shape(world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        v = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    A, B = workload_pixels(rank, world)
    npix = A.size
    rows_local = npix // cols
    npts_rank = npix * NGEO
    mode, _ = _engine.setup_mode(SPIN, th_o, A, B, True)
    obs = _engine.observer_struct(SPIN, R_O, th_o)
    d_alpha, d_beta = _engine.to_device(A, dev), _engine.to_device(B, dev)
    d_w = torch.empty((int(L.kgeo_ray_workspace_doubles(npix)),), dtype=torch.float64, device=dev)
    d_flags = torch.empty((npix,), dtype=torch.int32, device=dev)
    outs = [torch.empty((NGEO, npix), dtype=torch.float64, device=dev) for _ in range(4)]    # t, r, th, ph
    null = ctypes.cast(None, _lib.c_double_p)
    stream = _engine.stream_ptr()

    # image of the same pixels (N > 1: part of the step; N = 1: the secondary metric below uses its own 4096^2 plane)
    with_image = world > 1
    peer = None
    if with_image:
        iobs = _engine.observer_struct(SPIN, float("inf"), th_o)
        M = pack_models(ei.emis_default, ei.vel_default, ei.bfield_default)
        iopts = _lib.ImageOpts(0, 2, 0, 0, float(ei.SPECIND), ei.DISKANGLE, ei.OBSFREQ)
        npix_all = rows * cols
        # rank 0's planes are a peer buffer every rank writes over NVLink, one of two ways (KGEO_BENCH_GATHER):
        #   copy_engine  the image kernel writes local planes, then ONE strided 3-D copy by the copy engines on a side stream puts
        #                them at their rows of rank 0's planes while setup and the track pass run (no SM is taken from them);
        #   peer_stores  the image kernel stores every result straight to rank 0's planes (kgeo_equatorial_image_sharded) on a
        #                side stream beside setup and tracks.
        # None: the GPUs cannot map each other (or KGEO_B200_GATHER=nccl) -> all_gather_into_tensor + un-interleave on rank 0.
        peer = sharding.peer_buffer(IMG_PLANES * npix_all, 0)
        gather_mode = os.environ.get("KGEO_BENCH_GATHER", GATHER_DEFAULT) if peer is not None else "nccl_all_gather"
        if gather_mode not in ("copy_engine", "peer_stores", "nccl_all_gather"):
            raise ValueError("KGEO_BENCH_GATHER must be copy_engine or peer_stores")
        side = torch.cuda.Stream(device=dev, priority=-1)
        side_ptr = ctypes.c_void_p(side.cuda_stream)
        evi = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev_img = torch.cuda.Event()
        token = torch.zeros((1,), dtype=torch.float32, device=dev)
        img = gathered = None
        if peer is not None:
            smap = _lib.ShardMap(cols, world, rank, npix_all, 0)
            vp = lambda plane: ctypes.cast(ctypes.c_void_p(peer.ptr + 8 * plane * npix_all), _lib.c_double_p)
            peer_ptrs = [vp(3 * i) for i in range(7)] + [vp(21 + i) for i in range(3)]
            final = peer.tensor((IMG_PLANES, npix_all))
        if gather_mode != "peer_stores":
            img = torch.empty((IMG_PLANES, npix), dtype=torch.float64, device=dev)        # this rank's rows, plane-major
            img_ptrs = [_engine.dptr(img[3 * i:3 * i + 3]) for i in range(7)] + [_engine.dptr(img[21 + i]) for i in range(3)]
        if gather_mode == "nccl_all_gather":
            gathered = torch.empty((world, IMG_PLANES, npix), dtype=torch.float64, device=dev)
            final = torch.empty((IMG_PLANES, npix_all), dtype=torch.float64, device=dev) if rank == 0 else None

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    main = torch.cuda.current_stream(dev)

    def image_local():
        _lib.check(L.kgeo_equatorial_image(ctypes.byref(iobs), ctypes.byref(M), ctypes.byref(iopts), npix,
                                           _engine.dptr(d_alpha), _engine.dptr(d_beta), *img_ptrs, null, null, stream))

    def step(timed=False):
        mode_g = gather_mode if with_image else None
        if mode_g == "peer_stores":                  # image + peer-store gather on the side stream, beside setup and tracks
            side.wait_stream(main)
            if timed:
                evi[0].record(side)
            _lib.check(L.kgeo_equatorial_image_sharded(ctypes.byref(iobs), ctypes.byref(M), ctypes.byref(iopts), npix,
                                                       _engine.dptr(d_alpha), _engine.dptr(d_beta), ctypes.byref(smap),
                                                       *peer_ptrs, null, null, side_ptr))
            if timed:
                evi[1].record(side)
        elif mode_g == "copy_engine":                # image first, its planes travel while setup and tracks compute
            if timed:
                evi[0].record()
            image_local()
            if timed:
                evi[1].record()
            ev_img.record()
            side.wait_event(ev_img)
            _lib.check(L.kgeo_shard_scatter(_engine.dptr(img), IMG_PLANES, npix, ctypes.byref(smap),
                                            ctypes.cast(ctypes.c_void_p(peer.ptr), _lib.c_double_p), side_ptr))
        if timed:
            ev[0].record()
        _lib.check(L.kgeo_ray_setup(ctypes.byref(obs), npix, _engine.dptr(d_alpha), _engine.dptr(d_beta), mode,
                                    _engine.dptr(d_w), _engine.iptr(d_flags), stream))
        if timed:
            ev[1].record()
        _lib.check(L.kgeo_tracks(ctypes.byref(obs), npix, _engine.dptr(d_w), _engine.iptr(d_flags), NGEO, MAXTAUFRAC, null,
                                 0, 1, _engine.dptr(outs[0]), _engine.dptr(outs[1]), _engine.dptr(outs[2]),
                                 _engine.dptr(outs[3]), null, null, npix, stream))
        if timed:
            ev[2].record()
        if mode_g in ("peer_stores", "copy_engine"):
            main.wait_stream(side)
            dist.all_reduce(token)                   # completion signal: behind every rank's writes in stream order
            if timed:
                ev[4].record()
        elif with_image:
            image_local()
            if timed:
                ev[3].record()
            sharding.gather_interleaved_rows(img, gathered, final, rows_local, cols)
            if timed:
                ev[4].record()

    for _ in range(args.warmup):
        step()
    # per-phase durations (CUDA events on the launching stream), measured outside the headline region
    setup_ms, track_ms, image_ms, coll_ms = [], [], [], []
    for _ in range(3):
        barrier()
        step(timed=True)
        torch.cuda.synchronize()
        setup_ms.append(ev[0].elapsed_time(ev[1]))
        track_ms.append(ev[1].elapsed_time(ev[2]))
        if with_image and gather_mode != "nccl_all_gather":
            image_ms.append(evi[0].elapsed_time(evi[1]))      # the image kernel (peer_stores: on its own stream beside the tracks)
            coll_ms.append(ev[2].elapsed_time(ev[4]))         # exposed: end of the tracks -> planes complete on rank 0
        elif with_image:
            image_ms.append(ev[2].elapsed_time(ev[3]))
            coll_ms.append(ev[3].elapsed_time(ev[4]))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = _lib.launch_count()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step()
    t_stop.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_per_step = allmax(t_start.elapsed_time(t_stop)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * npts_rank / (ms_per_step * 1e-3)
    phases = {"setup_ms": allmax(np.median(setup_ms)), "track_ms": allmax(np.median(track_ms))}
    if with_image:
        phases["image_ms"] = allmax(np.median(image_ms))
        phases["collective_ms"] = allmax(np.median(coll_ms))

    if profile_only:                                  # under ncu: kernels only, no e2e / CPU legs, no JSON claim
        if rank == 0:
            emit({"profile_only": True, "ms_per_step": ms_per_step, "track_kernel_ms": float(np.median(track_ms)),
                  "setup_kernel_ms": float(np.median(setup_ms)), "phases": phases,
                  "gather": gather_mode if with_image else None})
        if world > 1:
            sharding.release_peer_buffers()
            dist.destroy_process_group()
        return

    # ---- the gathered planes on rank 0 against the same rows computed locally (bit for bit) ----
    gather_check = None
    if with_image:
        ok = 1
        if rank == 0:
            probe = sorted(set([1, rows // 2 + 1, rows - 1]))                 # rows owned by other ranks
            xs, ys = plane_pixels(world)
            sel = np.concatenate([np.arange(r * cols, (r + 1) * cols) for r in probe])
            pa = _engine.to_device(np.tile(xs, len(probe)), dev)
            pb = _engine.to_device(np.repeat(ys[probe], cols), dev)
            chk = torch.empty((IMG_PLANES, sel.size), dtype=torch.float64, device=dev)
            cp = [_engine.dptr(chk[3 * i:3 * i + 3]) for i in range(7)] + [_engine.dptr(chk[21 + i]) for i in range(3)]
            _lib.check(L.kgeo_equatorial_image(ctypes.byref(iobs), ctypes.byref(M), ctypes.byref(iopts), sel.size,
                                               _engine.dptr(pa), _engine.dptr(pb), *cp, null, null, stream))
            got = final[:, torch.as_tensor(sel, device=dev)]
            same = (got == chk) | (torch.isnan(got) & torch.isnan(chk))
            ok = int(bool(same.all()))
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            raise RuntimeError("bench: the gathered image planes differ from the rows computed on rank 0")
        gather_check = "ok"

    # sanity of the produced data (finite radii on the first samples of every ray)
    if not bool(torch.isfinite(outs[1][1, ::997]).all()):
        raise RuntimeError("bench: non-finite radii in the output")

    # ---- parity of the benched launch ----
    parity, cpu_baseline = None, None
    if world == 1:
        # the oracle on PARITY_PIXELS cfg-3 pixels x 500 samples, 1 core: the timing is the cpu_baseline, the values the checker
        from oracle import gates
        cpu_value, cpu_dt, cpu_pts, idx, res = time_cpu(PARITY_PIXELS, 1, 1, 0, keep=True)
        ref = res[0]
        cpu_baseline = {"value": cpu_value, "unit": "points/s", "cores": 1, "kind": "port",
                        "sample": "%d pixels (spread over the cfg-3 grid) x %d samples = %d points, %.1f s, oracle port of "
                                  "the reference's SciPy path; host has %d cores" % (PARITY_PIXELS, NGEO, cpu_pts, cpu_dt,
                                                                                     os.cpu_count() or 1)}
        ix = torch.as_tensor(idx, device=dev)
        got = {k: outs[i][:, ix].cpu().numpy() for i, k in enumerate(("t", "r", "th", "ph"))}
        budget = gates.parity_budget(got, ref, ref["tau"], SPIN, R_O, th_o, A[idx], B[idx])
        parity = {"against": "oracle/kgeo_oracle.py (pinned to the unmodified reference by tests/golden), same pixels, ngeo=500",
                  "pixels": PARITY_PIXELS, "samples": budget["n_samples"], "n_pv_masked": budget["n_pv_masked"],
                  "n_tail": budget["n_tail"], "n_axis": budget["n_axis"], "passes_gates": gates.budget_passes(budget)}
        for k, c in budget["coords"].items():
            parity[k] = {"max_rel": c["plain"]["max"], "p99.9": c["plain"]["p999"], "frac_gt_1e-10": c["plain"]["frac_gt_1e10"],
                         "gated_max": c["gated"]["max"], "gated_frac_gt_1e-10": c["gated"]["frac_gt_1e10"],
                         "pv_max": c["pv_max"], "tail_max": c["tail_max"], "nan_pattern_equal": c["nan_pattern_equal"]}
        if not parity["passes_gates"]:
            emit({"error": "parity gate broken on the benched workload", "parity": parity})
            raise RuntimeError("bench: parity gate broken on the benched workload: %r" % parity)
    else:
        parity = {"shard_check": shard_check(world, rank, dev)}

    # ---- free the resident tracks before the host-buffer legs (N = 2 holds 134 GB of them) ----
    del outs
    if with_image:
        del gathered, final, img
        sharding.release_peer_buffers()
    torch.cuda.empty_cache()

    e2e = e2e_legs(L, obs, mode, A, B, npix, cols, world, rank, local, dev, barrier, allmax)

    sweep = cfg5_sweep(world, rank, dev, barrier, allmax) if (world == 8 or (world > 1 and os.environ.get("KGEO_BENCH_SWEEP"))) else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the track pass) ----
    flops, pms = ctypes.c_double(), ctypes.c_float()
    peak_tf = None
    try:
        best = 0.0
        for _ in range(3):
            _lib.check(L.kgeo_fp64_peak_probe(20000, ctypes.byref(flops), ctypes.byref(pms), stream))
            best = max(best, flops.value / (pms.value * 1e-3) / 1e12)
        peak_tf = best
    except Exception:
        peak_tf = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    summ = ncu_summary()
    trk = (summ or {}).get("track_pass", {})
    fpp = trk.get("flop_per_point")
    tk = float(np.median(track_ms)) * 1e-3
    ach_tf = fpp * npts_rank / tk / 1e12 if fpp else None
    canon_tf = CANONICAL_FLOP_PER_POINT * npts_rank / tk / 1e12
    nominal_tf = 148 * 64 * 2 * 1.965e9 / 1e12
    roofline = {
        "bound": "fp64", "kernel": "track_hot_kernel + track_kernel<full, fix-up> (one track pass)",
        "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (ach_tf / peak_tf) if (peak_tf and ach_tf) else None,
        "flop_per_point": fpp,
        "flop_source": ("executed FP64 flops per point, ncu thread-instruction counts (2 DFMA + DMUL + DADD) of the capture in "
                        "profiles/r02_ncu_summary.json (git %s)" % summ.get("git_sha", "?")) if summ else
                       "null: the kernel sources changed since the capture in profiles/r02_ncu_summary.json",
        "fp64_pipe_active_ncu": trk.get("fp64_pipe_active_hot"),
        "warp_instructions_per_32_points_ncu": trk.get("warp_inst_per_32_points_hot"),
        "canonical": {"flop_per_point": CANONICAL_FLOP_PER_POINT, "achieved": canon_tf,
                      "frac": (canon_tf / peak_tf) if peak_tf else None,
                      "note": "SURVEY 8(d) work model of the reference's algorithm (Carlson duplication per sample); this "
                              "kernel computes the same values with per-ray theta-function series, ~4x fewer flops"},
        "traffic": trk.get("dram_bytes_per_pass") if (world == 1 and summ) else None, "traffic_unit": "B/pass",
        "peak_source": "DFMA-chain probe (kgeo_fp64_peak_probe) measured live on this GPU; MEASURED_PEAKS.json has no "
                       "FP64 entry; nominal 148 SM x 64 FMA/clk x 1.965 GHz = %.1f TFLOP/s" % nominal_tf,
        "points_per_launch": npts_rank, "kernel_ms": tk * 1e3, "setup_kernel_ms": float(np.median(setup_ms)),
        "hbm": {"achieved": BYTES_PER_POINT * npts_rank / tk / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": BYTES_PER_POINT * npts_rank / tk / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
    }
    if world > 1 and ach_tf and peak_tf:
        # all ranks run the same kernels on interleaved rows: aggregate = N x this rank's rate against N x the measured peak
        step_tf = fpp * world * npts_rank / (ms_per_step * 1e-3) / 1e12
        roofline["aggregate"] = {"achieved_track_pass": ach_tf * world, "achieved_whole_step": step_tf, "peak": peak_tf * world,
                                 "frac_track_pass": ach_tf / peak_tf, "frac_whole_step": step_tf / (peak_tf * world),
                                 "note": "whole step = setup + tracks + image + gather; north_star target >= 0.30"}

    secondary = image_secondary(L, ei, pack_models, th_o, dev, stream, peak_tf, summ) if world == 1 else sweep
    if cpu_baseline is None:   # N > 1: a short 1-core run of the port on cfg-4 pixels
        cpu_value, cpu_dt, cpu_pts, _, _ = time_cpu(1024, 1, 1, 0, world)
        cpu_baseline = {"value": cpu_value, "unit": "points/s", "cores": 1, "kind": "port",
                        "sample": "1024 pixels (spread over the cfg-4 grid) x 500 samples = %d points, %.1f s, oracle port of the "
                                  "reference's SciPy path; host has %d cores" % (cpu_pts, cpu_dt, os.cpu_count() or 1)}

    bytes_tracks = 32 * npts_rank
    line = {
        "metric": "fp64 geodesic points/sec (raytrace_ana)", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(world), "points_per_gpu_per_step": npts_rank,
                   "outputs": "r,theta,phi,t fp64 [500, %d] = %.1f GB per GPU per step" % (npix, bytes_tracks / 1e9),
                   "l2": "each step writes %.1f GB per GPU (>> 126 MB L2) and re-reads no output; no flush needed" % (bytes_tracks / 1e9),
                   "parallelism": ("1 GPU" if world == 1 else "rows interleaved over %d ranks; no data-path collective in the "
                                   "trace; image planes (%.2f GB) to rank 0: %s"
                                   % (world, IMG_PLANES * 8 * rows * cols / 1e9,
                                      {"peer_stores": "stored in place by every rank's image kernel over NVLink peer memory, on a side "
                                                      "stream beside the track pass; one 4-byte all_reduce as the completion signal",
                                       "copy_engine": "computed into local planes, then put at their rows of rank 0's planes (peer memory "
                                                      "over NVLink) by the copy engines on a side stream while setup and the track pass "
                                                      "run; one 4-byte all_reduce as the completion signal"}.get(gather_mode)
                                      if gather_mode != "nccl_all_gather"
                                      else "all_gather_into_tensor + un-interleave on rank 0"))},
        "phases": phases, "roofline": roofline, "parity": parity, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks, "secondary": secondary,
    }
    if world > 1:
        line["image_ms"] = phases["image_ms"]
        line["collective_ms"] = phases["collective_ms"]
        line["shard_check"] = parity["shard_check"]
        line["gather"] = gather_mode
        line["gather_check"] = gather_check
        if gather_mode == "peer_stores":
            line["efficiency_limiter"] = ("step = setup %.1f + tracks %.1f (with the image kernel, %.1f ms on its own stream, sharing "
                                          "the SMs and storing to rank 0 over NVLink) + %.2f ms exposed wait for the planes (max over "
                                          "ranks); the image and the per-ray setup add time but no geodesic points" %
                                          (phases["setup_ms"], phases["track_ms"], phases["image_ms"], phases["collective_ms"]))
        elif gather_mode == "copy_engine":
            line["efficiency_limiter"] = ("step = image %.1f + setup %.1f + tracks %.1f ms, the planes travelling to rank 0 through the "
                                          "copy engines meanwhile, + %.2f ms from the end of the tracks to the completion signal (max "
                                          "over ranks: includes the skew between ranks); the image and the per-ray setup add time but "
                                          "no geodesic points" %
                                          (phases["image_ms"], phases["setup_ms"], phases["track_ms"], phases["collective_ms"]))
        else:
            line["efficiency_limiter"] = ("step = setup %.1f + tracks %.1f + image %.1f + gather %.1f ms (max over ranks); the image, "
                                          "its gather and the per-ray setup add time but no geodesic points" %
                                          (phases["setup_ms"], phases["track_ms"], phases["image_ms"], phases["collective_ms"]))
        dist.barrier()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def shard_check(world, rank, dev):
    """N > 1: a 120 x 120 x 50 trace and a 250 x 250 polarized n = 0..2 image, sharded over the ranks, against the same
    computed whole on rank 0: bit for bit (tests/dist_check.py's check, run here because the driver's 1-GPU test tier skips it)."""
    import torch
    import torch.distributed as dist
    import kgeo_b200 as kg
    from kgeo_b200 import sharding
    a, th_o = 0.94, 17 * np.pi / 180
    with contextlib.redirect_stdout(io.StringIO()):
        kw = dict(emissivity=kg.Emissivity('ring'), velocity=kg.Velocity('general', fac_subkep=0.7, beta_phi=0.7, beta_r=0.7),
                  bfield=kg.Bfield('bz_guess', C=1.), polarization=True, specind=1.)
        res = sharding.make_image_sharded(a, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 250, dst=0, **kw)
        x = np.linspace(-8, 8, 120)
        A, B = [v.flatten() for v in np.meshgrid(x, x)]
        geos, (r0, r1) = sharding.raytrace_sharded(a, [0, 1000., th_o, 0], [A, B], 120, ngeo=50)
        whole = kg.raytrace_ana(a, [0, 1000., th_o, 0], [A, B], ngeo=50, plotdata=False)
        ok = all(np.array_equal(w[:, r0 * 120:r1 * 120], p, equal_nan=True) for w, p in zip(whole.geo_coords, geos.geo_coords))
        if rank == 0:
            single = kg.make_image(a, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 250, **kw)
            ok = ok and all(g.shape == w.shape and np.array_equal(g, w, equal_nan=True) for g, w in zip(res, single))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        raise RuntimeError("bench: sharded results differ from the whole-plane results")
    return "ok"


def cfg5_sweep(world, rank, dev, barrier, allmax):
    """N = 8: BASELINE cfg 5, 64 (spin, inclination) configurations of 512 x 512 n = 0..2 images dealt round-robin to the ranks
    (one fused launch per image) and gathered to rank 0; wall time of the whole sweep including the gather and the copy of the
    planes to host memory on rank 0."""
    import torch
    from kgeo_b200 import sharding
    spins = np.linspace(0.0, 0.99, 8)
    incs = np.linspace(5.0, 85.0, 8) * np.pi / 180
    cfgs = [(float(a), float(t)) for a in spins for t in incs]

    def run():
        with contextlib.redirect_stdout(io.StringIO()):
            return sharding.make_image_sweep_sharded(cfgs, np.inf, 2, -10., 10., -10., 10., 20. / 512, dst=0)
    run()
    barrier()
    t0 = time.perf_counter()
    res = run()
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0)
    if rank == 0 and not all(np.isfinite(r[0][::4099, 0]).all() for r in res):
        raise RuntimeError("bench: non-finite intensities in the cfg-5 sweep")
    return {"metric": "BASELINE cfg 5 wall time: 64 (a, theta_o) configs x 512x512 px n=0..2 images over %d GPUs, gathered to rank 0 "
                      "and copied to host" % world, "value": dt, "unit": "s", "pixels_per_s": 64 * 512 * 512 / dt,
            "configs": len(cfgs)}


def e2e_legs(L, obs, mode, A, B, npix, cols, world, rank, local, dev, barrier, allmax):
    """End-to-end legs with HOST buffers: C ABI (kgeo_raytrace_host), Python drop-in (raytrace_ana(...).geo_coords), and the
    box's D2H ceiling for the same bytes on all ranks at once.  On several GPUs each rank runs a bounded block of its rows
    (at most the cfg-3 volume, 16.8 GB of results): the legs are host-link-bound, more rows only repeat the measurement."""
    import torch
    import kgeo_b200
    from kgeo_b200 import _engine, _lib
    null = ctypes.cast(None, _lib.c_double_p)
    pin_cpus = local_cpu_affinity(local)     # pinned result buffers on the GPU's own NUMA node (first touch)
    try:
        import psutil
        avail = psutil.virtual_memory().available / max(1, world)
    except Exception:
        avail = 0
    e_rows = min(npix // cols, (1024 * 1024) // cols)
    while 4 * 8 * e_rows * cols * NGEO * 1.3 > avail and e_rows > 8:
        e_rows //= 2
    e_npix = e_rows * cols
    hA, hB = np.ascontiguousarray(A[:e_npix]), np.ascontiguousarray(B[:e_npix])
    host = [torch.empty((NGEO, e_npix), dtype=torch.float64, pin_memory=True) for _ in range(4)]
    hp = [ctypes.cast(h.data_ptr(), _lib.c_double_p) for h in host]
    nbytes = 32 * e_npix * NGEO

    def c_step():
        _lib.check(L.kgeo_raytrace_host(ctypes.byref(obs), e_npix, _engine.hptr(hA), _engine.hptr(hB), mode, NGEO,
                                        MAXTAUFRAC, hp[0], hp[1], hp[2], hp[3], null, null, local))

    def timed(fn, steps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return allmax((time.perf_counter() - t0) / steps)

    e_steps = 3
    c_dt = timed(c_step, e_steps)
    if not np.isfinite(host[1][1, ::997].numpy()).all():
        raise RuntimeError("bench: non-finite radii in the e2e output")

    # ceiling: the same bytes as bare pinned D2H copies (device slab -> the same pinned arrays), all ranks at once
    slab = torch.empty((min(NGEO, 64), e_npix), dtype=torch.float64, device=dev)

    def copy_step():
        for h in host:
            for k0 in range(0, NGEO, slab.shape[0]):
                k1 = min(NGEO, k0 + slab.shape[0])
                h[k0:k1].copy_(slab[:k1 - k0], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ceil_dt = timed(copy_step, 2)
    del slab
    _lib.check(L.kgeo_trim_memory(local))

    # Python drop-in: raytrace_ana(...).geo_coords (device-resident tracks, then one overlapped pass into pinned NumPy arrays;
    # the pinned blocks of `host` go back to torch's caching host allocator and are reused by the result arrays)
    del host, hp
    th_o = INC_DEG * np.pi / 180

    def py_step():
        with contextlib.redirect_stdout(io.StringIO()):
            g = kgeo_b200.raytrace_ana(SPIN, [0, R_O, th_o, 0], [hA, hB], ngeo=NGEO, plotdata=False)
            t, r, th, ph = g.geo_coords
        if not np.isfinite(r[1, ::997]).all():
            raise RuntimeError("bench: non-finite radii in the Python e2e output")
        del g, t, r, th, ph

    py_dt = timed(py_step, 2)
    torch.cuda.empty_cache()
    return {"value": world * e_npix * NGEO / c_dt, "unit": "points/s", "h2d_bytes_per_step": 16 * e_npix,
            "d2h_bytes_per_step": nbytes, "steps": e_steps, "ms_per_step": c_dt * 1e3, "pixels_per_rank": e_npix,
            "host_cpu_affinity": pin_cpus, "d2h_gbs": world * nbytes / c_dt / 1e9,
            "ceiling_gbs": world * nbytes / ceil_dt / 1e9,
            "frac_of_ceiling": ceil_dt / c_dt,
            "ceiling_note": "the same bytes as bare cudaMemcpyAsync D2H into the same pinned arrays on all %d rank(s) at once "
                            "(64-sample slabs): what this box's host links carry; e2e / ceiling says how much of it the call "
                            "reaches" % world,
            "python": {"value": world * e_npix * NGEO / py_dt, "unit": "points/s", "ms_per_step": py_dt * 1e3,
                       "frac_of_c_abi": c_dt / py_dt,
                       "call": "kgeo_b200.raytrace_ana(a, observer, [alpha, beta], ngeo=500).geo_coords (NumPy arrays)"},
            "note": "kgeo_raytrace_host: host alpha/beta -> setup + track kernels in 256 MiB sample slabs overlapped with "
                    "D2H into pinned host r/theta/phi/t arrays; wall clock around the blocking C call, max over ranks"}


def image_secondary(L, ei, pack_models, th_o, dev, stream, peak_tf, summ):
    """N = 1: the fused n = 0..2 equatorial image kernel (BASELINE cfg 2 models at cfg 4 size): kernel-only rate with its
    roofline, e2e through kgeo_equatorial_image_host with host buffers, and the oracle port's make_image at 256^2 beside it."""
    import torch
    from kgeo_b200 import _engine, _lib
    null = ctypes.cast(None, _lib.c_double_p)
    try:
        side = 4096
        x = np.linspace(-HALF_FOV, HALF_FOV, side)
        IA, IB = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, x)]
        npx = side * side
        dIA, dIB = _engine.to_device(IA, dev), _engine.to_device(IB, dev)
        img = torch.empty((IMG_PLANES, npx), dtype=torch.float64, device=dev)
        ptrs = [_engine.dptr(img[3 * i:3 * i + 3]) for i in range(7)] + [_engine.dptr(img[21 + i]) for i in range(3)]
        iobs = _engine.observer_struct(SPIN, float("inf"), th_o)
        M = pack_models(ei.emis_default, ei.vel_default, ei.bfield_default)
        opts = _lib.ImageOpts(0, 2, 0, 0, float(ei.SPECIND), ei.DISKANGLE, ei.OBSFREQ)

        def run_img():
            _lib.check(L.kgeo_equatorial_image(ctypes.byref(iobs), ctypes.byref(M), ctypes.byref(opts), npx, _engine.dptr(dIA),
                                               _engine.dptr(dIB), *ptrs, null, null, stream))
        for _ in range(3):
            run_img()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(5):
            run_img()
        i1.record()
        torch.cuda.synchronize()
        img_ms = i0.elapsed_time(i1) / 5
        if not bool(torch.isfinite(img[0, ::9973]).all()):
            raise RuntimeError("non-finite image")
        # e2e: host alpha/beta in, all 24 planes back into pinned host memory, through the C ABI
        himg = torch.empty((IMG_PLANES, npx), dtype=torch.float64, pin_memory=True)
        hp = [ctypes.cast(himg[3 * i:3 * i + 3].data_ptr(), _lib.c_double_p) for i in range(7)] + \
             [ctypes.cast(himg[21 + i].data_ptr(), _lib.c_double_p) for i in range(3)]

        def e2e_img():
            _lib.check(L.kgeo_equatorial_image_host(ctypes.byref(iobs), ctypes.byref(M), ctypes.byref(opts), npx, _engine.hptr(IA),
                                                    _engine.hptr(IB), *hp, null, null, torch.cuda.current_device()))
        e2e_img()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_img()
        e_dt = (time.perf_counter() - t0) / 3
        same = bool(torch.equal(himg[0], img[0].cpu()))
        # CPU: the oracle port's make_image at 256 x 256 (BASELINE cfg 2), 1 core
        from oracle import kgeo_oracle as ko
        np.seterr(all="ignore")
        c0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ko.make_image(SPIN, np.inf, th_o, 2, -10., 10., -10., 10., 20. / 256, dict(type='bpl'), dict(type='zamo'),
                          dict(type='rad'))
        c_dt = time.perf_counter() - c0
        isum = (summ or {}).get("image_kernel", {})
        fpx = isum.get("flop_per_pixel")
        ach = fpx * npx / (img_ms * 1e-3) / 1e12 if fpx else None
        return {"metric": "equatorial image pixels/s (n=0..2, all orders in one call: geometry kernel + emission kernel)", "value": npx / (img_ms * 1e-3),
                "unit": "pixels/s", "workload": "4096x4096 px, a=0.94, theta_o=17deg, r_o=inf, bpl/zamo/rad (BASELINE cfg 2 "
                "models at cfg 4 size)", "ms": img_ms, "includes": "kgeo_equatorial_image with device pointers: r_equatorial_kernel + "
                "image_emission_kernel (outputs preallocated, resident)",
                "roofline": {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": (ach / peak_tf) if (ach and peak_tf) else None, "flop_per_pixel": fpx,
                             "flop_source": "executed FP64 flops per pixel (ncu, profiles/r02_ncu_summary.json)" if fpx else
                                            "null: kernel sources changed since the capture",
                             "fp64_pipe_active_ncu": isum.get("fp64_pipe_active"),
                             "canonical_frac": CANONICAL_FLOP_PER_PIXEL * npx / (img_ms * 1e-3) / 1e12 / peak_tf if peak_tf else None,
                             "traffic": isum.get("dram_bytes_per_launch"), "algorithmic_bytes": (16 + 8 * IMG_PLANES) * npx,
                             "hbm_gbs": (16 + 8 * IMG_PLANES) * npx / (img_ms * 1e-3) / 1e9},
                "e2e": {"value": npx / e_dt, "unit": "pixels/s", "ms": e_dt * 1e3, "h2d_bytes_per_step": 16 * npx,
                        "d2h_bytes_per_step": 8 * IMG_PLANES * npx, "matches_device_path": same,
                        "call": "kgeo_equatorial_image_host (host alpha/beta in, 24 planes into pinned host memory)"},
                "cpu_baseline": {"value": 65536 / c_dt, "unit": "pixels/s", "cores": 1, "kind": "port",
                                 "sample": "oracle port of make_image, 256 x 256 px, n = 0..2 (BASELINE cfg 2), %.2f s" % c_dt}}
    except Exception as exc:                                       # the headline metric must still be reported
        return {"error": repr(exc)}


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the real stdout (see main: libraries print to stderr meanwhile)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line: anything a library prints (e.g. "NCCL version ..." on communicator creation)
    # is sent to stderr by pointing fd 1 at fd 2 for the duration of the run
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
