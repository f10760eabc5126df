#!/bin/bash
# round 2, GPU session 7: per-ray fix-up + relaxed polar theta bounds; image kernel with one copy of the model code
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -4 gpurun_out/s7_pytest.log
KGEO_BENCH_PROFILE=1 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1
timeout 600 python tests/hostsim/variant_image_bench.py > gpurun_out/s7_image_variants.txt 2>&1
cat gpurun_out/s7_image_variants.txt
timeout 600 python tools/parity_budget.py > gpurun_out/s7_parity_budget.log 2>&1
tail -9 gpurun_out/s7_parity_budget.log
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"track_hot_kernel|track_kernel" -s 6 -c 2 -o gpurun_out/r02_track_pass python bench.py --steps 1 --warmup 3 > gpurun_out/s7_ncu_track.log 2>&1
PROFILE_ONCE=1 timeout 600 ncu --set full --clock-control none -k regex:"image_kernel|r_equatorial_kernel" -o gpurun_out/r02_image python tools/profile_image.py 1 > gpurun_out/s7_ncu_image.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
