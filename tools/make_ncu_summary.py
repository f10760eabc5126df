"""Developer tool: profiles/r02_ncu_summary.json from ncu reports captured on the GPU box.

    python tools/make_ncu_summary.py --track gpurun_out/r02_track_pass.ncu-rep [--image gpurun_out/r02_image.ncu-rep]

bench.py reads the ncu-derived constants of its roofline blocks (executed FP64 flops per point / pixel, FP64-pipe activity, DRAM
traffic) from this file and uses them only while the CUDA sources still hash to `csrc_sha256_16`, i.e. while the kernels are
the ones that were profiled.  Executed flops = 2 x DFMA + DMUL + DADD thread-level, predicated-on instruction counts
(smsp__sass_thread_inst_executed_op_*_pred_on.sum.per_cycle_elapsed x sm__cycles_elapsed.avg)."""
import argparse, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def f(row, key):
    return float(row[key].replace(",", ""))


def flops(row):
    cyc = f(row, "sm__cycles_elapsed.avg")
    n = {op: f(row, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op) * cyc for op in ("dfma", "dmul", "dadd")}
    return 2 * n["dfma"] + n["dmul"] + n["dadd"], n


def gbytes(row, key):
    v, unit = f(row, key), None
    return v   # the raw page reports these in the unit of the header row; callers pass files captured with Gbyte/Mbyte units


def dram(rep_rows_units, row):
    return None


ap = argparse.ArgumentParser()
ap.add_argument("--track")
ap.add_argument("--image")
ap.add_argument("--points", type=float, default=1024 * 1024 * 500)
ap.add_argument("--pixels", type=float, default=4096 * 4096)
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_ncu_summary.json"))
args = ap.parse_args()


def units_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[1]))


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


git = os.environ.get("KGEO_GIT_SHA") or subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
dirty = (not os.environ.get("KGEO_GIT_SHA")) and bool(subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "kgeo_b200/csrc"], capture_output=True, text=True).stdout.strip())
summ = {"what": "ncu --set full --clock-control none captures of `bench.py` (KGEO_BENCH_PROFILE=1) and tools/profile_image.py on one B200",
        "git_sha": git + ("+uncommitted kernel changes" if dirty else ""), "csrc_sha256_16": bench.kernel_source_hash()}
if os.path.exists(args.out):
    try:
        old = json.load(open(args.out))
        if old.get("csrc_sha256_16") == summ["csrc_sha256_16"]:
            for k in ("track_pass", "image_kernel", "r_equatorial_kernel", "setup"):
                if k in old:
                    summ[k] = old[k]
    except Exception:
        pass
if args.track:
    rows, U = raw(args.track), units_of(args.track)
    hot = next(r for r in rows if "track_hot_kernel" in r["Kernel Name"])
    fix = next(r for r in rows if "track_kernel" in r["Kernel Name"] and "track_hot" not in r["Kernel Name"])
    fh, nh = flops(hot)
    ff, nf = flops(fix)
    traffic = sum(to_bytes(f(r, k), U[k]) for r in (hot, fix) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    summ["track_pass"] = {
        "source": os.path.basename(args.track),
        "flop_per_point": (fh + ff) / args.points, "flop_per_point_hot": fh / args.points, "flop_per_point_fixup": ff / args.points,
        "thread_inst_per_point_hot": {k: v / args.points for k, v in nh.items()},
        "fp64_pipe_active_hot": f(hot, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100,
        "issue_active_hot": f(hot, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
        "warp_inst_per_32_points_hot": f(hot, "smsp__inst_executed.sum") * 32 / args.points,
        "hot_ms": f(hot, "gpu__time_duration.sum"), "fixup_ms": f(fix, "gpu__time_duration.sum"),
        "registers_hot": f(hot, "launch__registers_per_thread"), "warps_active_hot": f(hot, "sm__warps_active.avg.pct_of_peak_sustained_active") / 100,
        "dram_bytes_per_pass": traffic,
        "dram_read_hot": to_bytes(f(hot, "dram__bytes_read.sum"), U["dram__bytes_read.sum"]),
        "dram_write_hot": to_bytes(f(hot, "dram__bytes_write.sum"), U["dram__bytes_write.sum"]),
    }
if args.image:
    rows, U = raw(args.image), units_of(args.image)
    # an image = r_equatorial_kernel (geometry) + image_emission_kernel; tools/profile_image.py runs 256^2 and 4096^2, unpolarized
    # default models and polarized + path length, then r_equatorial alone
    ker = [r for r in rows if "image_emission_kernel" in r["Kernel Name"] or "r_equatorial_kernel" in r["Kernel Name"]]
    pairs = [(ker[i - 1], ker[i]) for i in range(1, len(ker)) if "image_emission_kernel" in ker[i]["Kernel Name"]
             and "r_equatorial_kernel" in ker[i - 1]["Kernel Name"]]
    big = [p for p in pairs if f(p[1], "launch__grid_size") > 10000]
    dram = lambda r: to_bytes(f(r, "dram__bytes_read.sum"), U["dram__bytes_read.sum"]) + to_bytes(f(r, "dram__bytes_write.sum"), U["dram__bytes_write.sum"])
    pipe = lambda r: f(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100
    ms = lambda r: f(r, "gpu__time_duration.sum")
    if big:
        g, e = big[0]  # 4096^2 unpolarized default models (the launches bench.py's secondary metric times)
        fl = flops(g)[0] + flops(e)[0]
        summ["image_kernel"] = {
            "source": os.path.basename(args.image), "kernels": "r_equatorial_kernel (geometry) + image_emission_kernel",
            "flop_per_pixel": fl / args.pixels,
            "fp64_pipe_active": (pipe(g) * ms(g) + pipe(e) * ms(e)) / (ms(g) + ms(e)),
            "issue_active": (f(g, "smsp__issue_active.avg.pct_of_peak_sustained_active") * ms(g) +
                             f(e, "smsp__issue_active.avg.pct_of_peak_sustained_active") * ms(e)) / (ms(g) + ms(e)) / 100,
            "ms": ms(g) + ms(e), "geometry_ms": ms(g), "emission_ms": ms(e),
            "registers": [f(g, "launch__registers_per_thread"), f(e, "launch__registers_per_thread")],
            "warp_inst_per_32_pixels": (f(g, "smsp__inst_executed.sum") + f(e, "smsp__inst_executed.sum")) * 32 / args.pixels,
            "dram_bytes_per_launch": dram(g) + dram(e),
            "launches": [{"grid": f(x, "launch__grid_size"), "kernel": x["Kernel Name"].split("(")[0], "ms": ms(x), "fp64_pipe_active": pipe(x)}
                         for pr in pairs for x in pr]}
    req = [r for r in ker if "r_equatorial_kernel" in r["Kernel Name"] and f(r, "launch__grid_size") > 10000]
    if req:
        r = req[-1]    # the stand-alone r_equatorial call (n = 0..2, with Nmax planes)
        fl, n = flops(r)
        summ["r_equatorial_kernel"] = {"flop_per_pixel": fl / args.pixels, "ms": ms(r), "fp64_pipe_active": pipe(r),
                                       "dram_bytes_per_launch": dram(r)}
json.dump(summ, open(args.out, "w"), indent=1)
print(json.dumps(summ, indent=1))
