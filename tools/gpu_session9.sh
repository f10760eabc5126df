#!/bin/bash
# round 2, GPU session 9: tests, LDG variants, compute-sanitizer over every kernel family (odd sizes), full bench N = 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
tail -4 gpurun_out/s9_pytest.log
timeout 600 python tools/variant_bench.py > gpurun_out/s9_variants.txt 2>&1
cat gpurun_out/s9_variants.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/hostsim/sanitize_case.py > gpurun_out/s9_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/s9_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/hostsim/sanitize_case.py > gpurun_out/s9_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/s9_racecheck.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench_1gpu.json 2> gpurun_out/s9_bench_1gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/s9_bench_1gpu.err | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/s9_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','phases')})
print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','d2h_gbs','ceiling_gbs','frac_of_ceiling')}, d['e2e']['python'])
print('secondary', json.dumps(d['secondary'])[:1500])
print('parity', json.dumps(d['parity'])[:1200])
PY
