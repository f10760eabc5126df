#!/usr/bin/env python
"""bench.py -- fp64 geodesic points/s of the analytic Kerr ray tracer (BASELINE.json metric).

One "step" = one pass of the hot path over the workload: per-ray setup kernel + fused track kernel producing
r, theta, phi, t for every (pixel, Mino-time sample) of BASELINE config 3 (1024 x 1024 pixels x 500 samples, a=0.94,
theta_o=17 deg, r_o=1000, +-10 M).  With N>1 ranks (one process per GPU under torchrun) every rank traces its own
1024 interleaved rows (row % N == rank) of a (1024*N) x 1024 image plane -- weak scaling, no data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      FP64-pipe roofline of the track pass (hot kernel + fix-up kernel, back to back on one stream): achieved =
                FP64 flops the two kernels execute per point (608: the DFMA/DMUL/DADD mix of the committed ncu capture,
                profiles/) x points per launch / CUDA-event duration of the pass;
                peak = DFMA-chain rate measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry).  "canonical"
                restates the same launch in SURVEY 8(d)'s work model (2400 flop/point for Carlson duplication): the kernel
                evaluates the integrals by per-ray theta-function series instead, so that figure may exceed the peak.
                Plus the HBM side (32 B/point written) against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle port of the reference (oracle/kgeo_oracle.py, SciPy path) timed on this box's host cores on
                a bounded sample of the same workload.
  e2e           same metric through the host-buffer C-ABI call kgeo_raytrace_host: host alpha/beta in, r/theta/phi/t
                copied back into pinned host arrays inside the timed region.
"""
import argparse
import ctypes
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
NPIX_SIDE, NGEO = 1024, 500
MAXTAUFRAC = 1.0 - 1.0e-10
CANONICAL_FLOP_PER_POINT = 2400.0   # SURVEY.md 8(d) work model: Carlson duplication, the reference's algorithm
# FP64 flops the track pass executes per point on this workload (thread-level, predicated-on, ncu --set full of this
# command, profiles/r01_final4_ncu_full_subset.csv): hot kernel 223.9 DFMA x 2 + 113.2 DMUL + 35.0 DADD = 596.1, fix-up
# kernel (the 0.3 % Carlson-fallback rays, spread over all points) 12.3
EXECUTED_FLOP_PER_POINT = 608.0
FP64_PIPE_ACTIVE_NCU = 0.609        # sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active of the hot kernel, same capture
BYTES_PER_POINT = 32.0      # r, theta, phi, t written once


def workload_pixels(rank, world, side=NPIX_SIDE):
    """Rows rank, rank+world, ... of the (side*world) x side plane spanning +-10 M (cfg 3 when world == 1).  Rows are
    interleaved so that every rank sees the same mix of root classes: rays through the shadow cost more than
    rays that escape, and contiguous row blocks would make the central ranks the stragglers."""
    x = np.linspace(-HALF_FOV, HALF_FOV, side)
    y_all = np.linspace(-HALF_FOV, HALF_FOV, side * world)
    y = y_all[rank::world]
    A, B = np.meshgrid(x, y)
    return np.ascontiguousarray(A.flatten()), np.ascontiguousarray(B.flatten())


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's SciPy path (test infrastructure, used here only as the thing compared to)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    a, r_o, th_o, A, B, ngeo = args
    np.seterr(all="ignore")
    from oracle import kgeo_oracle as ko
    tau, sig, (t, r, th, ph) = ko.raytrace(a, [0, r_o, th_o, 0], [A, B], ngeo, MAXTAUFRAC)
    return float(np.nansum(r[1]))


def cpu_sample(npix_per_worker, workers):
    """A strided (class-representative) subset of the cfg-3 pixel grid, split into one chunk per worker."""
    A, B = workload_pixels(0, 1)
    n = npix_per_worker * workers
    stride = max(1, A.size // n)
    idx = np.arange(0, A.size, stride)[:n]
    th_o = INC_DEG * np.pi / 180
    return [(SPIN, R_O, th_o, A[idx[w::workers]], B[idx[w::workers]], NGEO) for w in range(workers)]


def time_cpu(npix_per_worker, workers, steps, warmup):
    chunks = cpu_sample(npix_per_worker, workers)
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
            run()
        dt = (time.perf_counter() - t0) / steps
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    return npts / dt, dt, npts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_worker = 512           # 512 px x 500 samples = 2.6e5 points per worker per step (~2-3 s of SciPy)
    value, dt, npts = time_cpu(per_worker, workers, args.steps, args.warmup)
    sample = "%d pixels (strided subset of the 1024x1024 cfg-3 grid) x %d samples per step, %d processes" % (
        per_worker * workers, NGEO, workers)
    line = {
        "impl": "reference", "metric": "fp64 geodesic points/sec (raytrace_ana)", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE cfg 3 geodesic tracks: a=0.94, theta_o=17deg, r_o=1000, 1024x1024 px x 500 samples",
                   "note": "reference algorithm (oracle port of kgeo's SciPy path; the Python reference itself cannot "
                           "travel to the GPU box) on the host cores"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler
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
    from kgeo_b200 import _engine, _lib

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

    A, B = workload_pixels(rank, world)
    npix = A.size
    npts_rank = npix * NGEO
    mode, _ = _engine.setup_mode(SPIN, th_o, A, B, True)
    obs = _engine.observer_struct(SPIN, R_O, th_o)
    d_alpha, d_beta = _engine.to_device(A, dev), _engine.to_device(B, dev)
    nw = L.kgeo_ray_workspace_slots()
    d_w = torch.empty((int(L.kgeo_ray_workspace_doubles(npix)),), dtype=torch.float64, device=dev)
    d_flags = torch.empty((npix,), dtype=torch.int32, device=dev)
    outs = [torch.empty((NGEO, npix), dtype=torch.float64, device=dev) for _ in range(4)]    # t, r, th, ph
    null = ctypes.cast(None, _lib.c_double_p)
    stream = _engine.stream_ptr()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    setup_ms, track_ms = [], []

    def step(timed=False):
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # per-kernel durations (CUDA events on the launching stream), measured outside the headline region
    for _ in range(3):
        step(timed=True)
        torch.cuda.synchronize()
        setup_ms.append(ev[0].elapsed_time(ev[1]))
        track_ms.append(ev[1].elapsed_time(ev[2]))

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
    ms = t_start.elapsed_time(t_stop)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tms.item()) / args.steps
    clocks = sampler.stop() if (rank == 0 and not os.environ.get("KGEO_BENCH_PROFILE")) else None
    value = world * npts_rank / (ms_per_step * 1e-3)

    # sanity of the produced data (finite radii on the first samples of every ray)
    chk = outs[1][1, ::997]
    if not bool(torch.isfinite(chk).all()):
        raise RuntimeError("bench: non-finite radii in the output")

    if os.environ.get("KGEO_BENCH_PROFILE"):      # under ncu: kernels only, no e2e / CPU legs, no JSON claim
        if rank == 0:
            sampler.stop()
            emit({"profile_only": True, "ms_per_step": ms_per_step, "track_kernel_ms": float(np.median(track_ms))})
        return

    # ---- e2e through the host-buffer C ABI ----
    e2e = None
    pin_cpus = local_cpu_affinity(local)     # pinned result buffers on the GPU's own NUMA node (first touch)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    need = 4 * 8 * npts_rank
    rows = NPIX_SIDE
    while need * 1.5 > avail and rows > 32:
        rows //= 2
        need = 4 * 8 * rows * NPIX_SIDE * NGEO
    e_npix = rows * NPIX_SIDE
    hA, hB = np.ascontiguousarray(A[:e_npix]), np.ascontiguousarray(B[:e_npix])
    host = [torch.empty((NGEO, e_npix), dtype=torch.float64, pin_memory=True) for _ in range(4)]
    hp = [ctypes.cast(h.data_ptr(), _lib.c_double_p) for h in host]

    def e2e_step():
        _lib.check(L.kgeo_raytrace_host(ctypes.byref(obs), e_npix, _engine.hptr(hA), _engine.hptr(hB), mode, NGEO,
                                        MAXTAUFRAC, hp[0], hp[1], hp[2], hp[3], null, null, local))

    e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_dt = (time.perf_counter() - t0) / e_steps
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": world * e_npix * NGEO / float(te.item()), "unit": "points/s", "h2d_bytes_per_step": 16 * e_npix,
           "d2h_bytes_per_step": 32 * e_npix * NGEO, "steps": e_steps, "ms_per_step": float(te.item()) * 1e3,
           "pixels_per_rank": e_npix, "host_cpu_affinity": pin_cpus,
           "note": "kgeo_raytrace_host: host alpha/beta -> setup + track kernels in 256 MiB sample slabs overlapped with "
                   "D2H into pinned host r/theta/phi/t arrays; wall clock around the blocking C call"}
    if not np.isfinite(host[1][1, ::997].numpy()).all():
        raise RuntimeError("bench: non-finite radii in the e2e output")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (track_kernel) ----
    flops = ctypes.c_double()
    pms = ctypes.c_float()
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
    tk = float(np.median(track_ms)) * 1e-3
    ach_tf = EXECUTED_FLOP_PER_POINT * npts_rank / tk / 1e12
    canon_tf = CANONICAL_FLOP_PER_POINT * npts_rank / tk / 1e12
    nominal_tf = 148 * 64 * 2 * 1.965e9 / 1e12
    roofline = {
        "bound": "fp64", "kernel": "track_kernel<full, hot> + track_kernel<full, fix-up> (one track pass)",
        "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (ach_tf / peak_tf) if peak_tf else None,
        "flop_per_point": EXECUTED_FLOP_PER_POINT,
        "flop_source": "FP64 flops executed per point (ncu: hot kernel 224 DFMA + 113 DMUL + 35 DADD thread instructions, "
                       "fix-up kernel 12 flop)",
        "fp64_pipe_active_ncu": FP64_PIPE_ACTIVE_NCU,
        "canonical": {"flop_per_point": CANONICAL_FLOP_PER_POINT, "achieved": canon_tf,
                      "frac": (canon_tf / peak_tf) if peak_tf else None,
                      "note": "SURVEY 8(d) work model of the reference's algorithm (Carlson duplication per sample); this "
                              "kernel computes the same values with per-ray theta-function series, ~4x fewer flops"},
        # dram__bytes_read.sum + dram__bytes_write.sum of one track pass on this workload, from the committed ncu --set full
        # capture (profiles/r01_final4_ncu_full_subset.csv): hot 2.19 + 16.92 GB, fix-up 0.39 + 0.45 GB; algorithmic 16.78 GB
        # of results + 1.62 GB of ray constants (2.16 GB workspace less the AGM ladders the hot pass does not stage)
        "traffic": 19.95e9 if npix == NPIX_SIDE * NPIX_SIDE else None, "traffic_unit": "B/pass",
        "peak_source": "DFMA-chain probe (kgeo_fp64_peak_probe) measured live on this GPU; MEASURED_PEAKS.json has no "
                       "FP64 entry; nominal 148 SM x 64 FMA/clk x 1.965 GHz = %.1f TFLOP/s" % nominal_tf,
        "points_per_launch": npts_rank, "kernel_ms": tk * 1e3,
        "setup_kernel_ms": float(np.median(setup_ms)),
        "hbm": {"achieved": BYTES_PER_POINT * npts_rank / tk / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": BYTES_PER_POINT * npts_rank / tk / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
    }

    # ---- secondary: fused n=0..2 equatorial image (BASELINE cfg 2 models at 4096 x 4096), kernel only ----
    secondary = None
    try:
        from kgeo_b200 import equatorial_images as ei
        from kgeo_b200.models import pack_models
        side = 4096
        x = np.linspace(-HALF_FOV, HALF_FOV, side)
        IA, IB = [np.ascontiguousarray(v.flatten()) for v in np.meshgrid(x, x)]
        npx = side * side
        dIA, dIB = _engine.to_device(IA, dev), _engine.to_device(IB, dev)
        planes = [torch.empty((3, npx), dtype=torch.float64, device=dev) for _ in range(7)]
        scal = [torch.empty((npx,), dtype=torch.float64, device=dev) for _ in range(3)]
        iobs = _engine.observer_struct(SPIN, float("inf"), th_o)
        M = pack_models(ei.emis_default, ei.vel_default, ei.bfield_default)
        opts = _lib.ImageOpts(0, 2, 0, 0, float(ei.SPECIND), ei.DISKANGLE, ei.OBSFREQ)

        def run_img():
            _lib.check(L.kgeo_equatorial_image(ctypes.byref(iobs), ctypes.byref(M), ctypes.byref(opts), npx, _engine.dptr(dIA),
                                               _engine.dptr(dIB), *[_engine.dptr(p) for p in planes],
                                               *[_engine.dptr(v) for v in scal], null, null, stream))
        run_img()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(5):
            run_img()
        i1.record()
        torch.cuda.synchronize()
        img_ms = i0.elapsed_time(i1) / 5
        if not bool(torch.isfinite(planes[0][0, ::9973]).all()):
            raise RuntimeError("non-finite image")
        secondary = {"metric": "equatorial image pixels/s (n=0..2, all orders fused)", "value": npx / (img_ms * 1e-3),
                     "unit": "pixels/s", "workload": "4096x4096 px, a=0.94, theta_o=17deg, r_o=inf, bpl/zamo/rad (BASELINE cfg 2 "
                     "models at cfg 4 size)", "ms": img_ms, "includes": "image_kernel only (outputs preallocated, resident)",
                     "fp64_frac": 3000.0 * npx / (img_ms * 1e-3) / 1e12 / peak_tf if peak_tf else None,
                     "hbm_gbs": 184.0 * npx / (img_ms * 1e-3) / 1e9}
        del planes, scal, dIA, dIB
    except Exception as exc:                                       # the headline metric must still be reported
        secondary = {"error": repr(exc)}

    # ---- CPU baseline on this box's host cores (bounded sample, 1 core: the reference as shipped is single-threaded) ----
    cpu_value, cpu_dt, cpu_pts = time_cpu(4096, 1, 1, 0)
    cpu_baseline = {"value": cpu_value, "unit": "points/s", "cores": 1, "kind": "port",
                    "sample": "4096 pixels (strided subset of the cfg-3 grid) x 500 samples = %d points, %.1f s, "
                              "oracle port of the reference's SciPy path; host has %d cores" % (cpu_pts, cpu_dt, os.cpu_count() or 1)}

    line = {
        "metric": "fp64 geodesic points/sec (raytrace_ana)", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE cfg 3 geodesic tracks: a=0.94, theta_o=17deg, r_o=1000, +-10M, 1024x1024 px x 500 "
                               "Mino-time samples -> r,theta,phi,t (per GPU; N GPUs trace a (1024N)x1024 plane by rows)",
                   "points_per_gpu_per_step": npts_rank, "outputs": "r,theta,phi,t fp64 [500, 1048576] = 16.8 GB per step",
                   "l2": "each step writes 16.8 GB (>> 126 MB L2) and re-reads no output; no flush needed",
                   "parallelism": "rows interleaved over %d rank(s), no collective" % world},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "secondary": secondary,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
