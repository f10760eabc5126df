#!/bin/bash
# round 2, GPU session 27 (1 GPU): image split into geometry + emission kernels, harmonic cut-off 1e-17: GPU suite, image launch shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s27_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s27_pytest.log | cut -c1-250
timeout 600 python tools/image_bench.py > gpurun_out/s27_image_variants.txt 2>&1; cat gpurun_out/s27_image_variants.txt
KGEO_BENCH_PROFILE=1 timeout 300 python bench.py --steps 5 --warmup 3 | tail -1 | cut -c1-300
