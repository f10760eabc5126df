"""Developer tool (GPU): time bench.py's kernel-only loop for every library build under kgeo_b200/lib/variants/."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for so in sorted(glob.glob(os.path.join(ROOT, "kgeo_b200", "lib", "variants", "*.so"))):
    env = dict(os.environ, KGEO_B200_LIB=so, KGEO_BENCH_PROFILE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3"], env=env,
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("%-28s step %.2f ms  track kernel %.2f ms" % (os.path.basename(so), d["ms_per_step"], d["track_kernel_ms"]), flush=True)
    except Exception:
        print(os.path.basename(so), "FAILED", out.stderr[-300:], flush=True)
