#!/bin/bash
# round 2, GPU session 17 (1 GPU): lazy Geodesics (geo_coords through the host pipeline): GPU suite + bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s17_pytest.log
timeout 900 python bench.py > gpurun_out/s17_bench_1gpu.json 2> gpurun_out/s17_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/s17_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s17_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])
print('e2e', {k:d['e2e'][k] for k in ('value','d2h_gbs','ceiling_gbs','frac_of_ceiling')}, d['e2e']['python'])
PY
