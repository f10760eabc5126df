#!/bin/bash
# round 2, GPU session 10: jet FF/gammamax tests, image kernel (two-loop, POL template) launch-bound sweep, setup with single
# evaluation of the complete integrals
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10_pytest.log
tail -4 gpurun_out/s10_pytest.log
KGEO_BENCH_PROFILE=1 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1
timeout 600 python tools/image_bench.py > gpurun_out/s10_image_variants.txt 2>&1
cat gpurun_out/s10_image_variants.txt
