#!/bin/bash
# round 2, GPU session 3: TMA-fed hot kernel on the grouped workspace: full GPU tests, A/B variants, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -5 gpurun_out/s3_pytest.log
timeout 600 python tools/variant_bench.py > gpurun_out/s3_variants.txt 2>&1
cat gpurun_out/s3_variants.txt
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_hot_kernel -s 3 -c 1 -o gpurun_out/r02_hot_tma python bench.py --steps 1 --warmup 3 > gpurun_out/s3_ncu_hot.log 2>&1
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ray_setup_kernel|theta_setup_kernel" -s 6 -c 2 -o gpurun_out/r02_setup_base python bench.py --steps 1 --warmup 3 > gpurun_out/s3_ncu_setup.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
