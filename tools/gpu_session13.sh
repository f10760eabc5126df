#!/bin/bash
# round 2, GPU session 13: one-logarithm R1 / S1 on plunging rays: tests, timing, parity budget, ncu of the track pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13_pytest.log
tail -4 gpurun_out/s13_pytest.log; grep -n "^E " gpurun_out/s13_pytest.log | head -5
KGEO_BENCH_PROFILE=1 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1
timeout 600 python tools/parity_budget.py > gpurun_out/s13_parity_budget.log 2>&1
tail -9 gpurun_out/s13_parity_budget.log | cut -c1-330
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"track_hot_kernel|track_kernel" -s 6 -c 2 -o gpurun_out/r02_track_pass python bench.py --steps 1 --warmup 3 > gpurun_out/s13_ncu_track.log 2>&1
ls -la gpurun_out | head
