#!/bin/bash
# round 2, GPU session 12: setup with the shared complete-integral bundle: tests, timings, parity budget
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12_pytest.log
tail -4 gpurun_out/s12_pytest.log
KGEO_BENCH_PROFILE=1 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1
timeout 600 python tools/image_bench.py > gpurun_out/s12_image.txt 2>&1
cat gpurun_out/s12_image.txt
timeout 600 python tools/parity_budget.py > gpurun_out/s12_parity_budget.log 2>&1
tail -9 gpurun_out/s12_parity_budget.log | cut -c1-330
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ray_setup_kernel|theta_setup_kernel" -s 6 -c 2 -o gpurun_out/r02_setup python bench.py --steps 1 --warmup 3 > gpurun_out/s12_ncu_setup.log 2>&1
ls -la gpurun_out | head -20
