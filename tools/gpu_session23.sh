#!/bin/bash
# round 2, GPU session 23: final captures -> profiles/r02_ncu_summary.json, then the bench lines that read it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s23_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s23_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/hostsim/sanitize_case.py > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_compute_sanitizer_memcheck.log
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"track_hot_kernel|track_kernel" -s 6 -c 2 -o gpurun_out/r02_track_pass python bench.py --steps 1 --warmup 3 > gpurun_out/s23_ncu_track.log 2>&1
PROFILE_ONCE=1 timeout 600 ncu --set full --clock-control none -k regex:"image_emission_kernel|r_equatorial_kernel" -o /tmp/r02_image python tools/profile_image.py 1 > gpurun_out/s23_ncu_image.log 2>&1
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none -k regex:"ray_setup_kernel|theta_setup_kernel" -s 9 -c 3 -o /tmp/r02_setup python bench.py --steps 1 --warmup 3 > gpurun_out/s23_ncu_setup.log 2>&1
KGEO_GIT_SHA=0f98296 python tools/make_ncu_summary.py --track gpurun_out/r02_track_pass.ncu-rep --image /tmp/r02_image.ncu-rep > gpurun_out/s23_summary.log 2>&1
cp profiles/r02_ncu_summary.json gpurun_out/r02_ncu_summary.json
ncu -i /tmp/r02_image.ncu-rep --page raw --csv > gpurun_out/r02_image_raw.csv 2>/dev/null
ncu -i /tmp/r02_setup.ncu-rep --page raw --csv > gpurun_out/r02_setup_raw.csv 2>/dev/null
KGEO_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/s23_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/s23_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','phases')}); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['flop_per_point'], d['roofline']['fp64_pipe_active_ncu'])
print(d['secondary']['value'], d['secondary']['roofline'])
r=json.load(open('gpurun_out/r02_bench_reference_arm.json')); print('ref', r['value'], r['cpu_baseline'])
PY
ls -la gpurun_out | head -20
