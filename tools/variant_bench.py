"""Developer tool (GPU): time bench.py's kernel-only loop for the default library (also with the round-1 hot pass,
KGEO_B200_TRK_PASSES=oldhot) and for every library build under kgeo_b200/lib/variants/ (tests/hostsim/build_variants.sh)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
runs = [("default", {}), ("default+oldhot", {"KGEO_B200_TRK_PASSES": "oldhot"})]
runs += [(os.path.basename(so), {"KGEO_B200_LIB": so}) for so in sorted(glob.glob(os.path.join(ROOT, "kgeo_b200", "lib", "variants", "*.so")))]
for name, extra in runs:
    env = dict(os.environ, KGEO_BENCH_PROFILE="1", **extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3"], env=env,
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("%-28s step %.2f ms  track pass %.2f ms  setup %.3f ms" % (name, d["ms_per_step"], d["track_kernel_ms"], d["setup_kernel_ms"]), flush=True)
    except Exception:
        print(name, "FAILED", out.stderr[-300:], flush=True)
