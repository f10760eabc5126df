#!/bin/bash
# round 2, GPU session 6: group-level hot / fix-up split, sincos by value: tests, variants, parity budget, ncu of the track pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
tail -4 gpurun_out/s6_pytest.log
timeout 600 python tools/variant_bench.py > gpurun_out/s6_variants.txt 2>&1
cat gpurun_out/s6_variants.txt
timeout 900 python tools/parity_budget.py > gpurun_out/s6_parity_budget.log 2>&1
cat gpurun_out/s6_parity_budget.log | tail -12
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"track_hot_kernel|track_kernel" -s 6 -c 2 -o gpurun_out/r02_track_pass python bench.py --steps 1 --warmup 3 > gpurun_out/s6_ncu_track.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
