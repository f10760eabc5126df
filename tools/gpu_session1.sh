#!/bin/bash
# round 2, GPU session 1: tests with the new hot kernel, variant timings, ncu baselines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
timeout 600 python tools/variant_bench.py > gpurun_out/s1_variants.txt 2>&1
cat gpurun_out/s1_variants.txt
timeout 300 python tools/profile_image.py 5 > gpurun_out/s1_image_times.txt 2>&1
cat gpurun_out/s1_image_times.txt
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_hot_kernel -s 3 -c 1 -o gpurun_out/r02_hot_ns2 python bench.py --steps 1 --warmup 3 > gpurun_out/s1_ncu_hot.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"image_kernel|r_equatorial_kernel" -o gpurun_out/r02_image_base python tools/profile_image.py 1 > gpurun_out/s1_ncu_image.log 2>&1
ls -la gpurun_out
