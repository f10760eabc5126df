#!/bin/bash
# round 2, GPU session 31 (1 GPU): per-ray setup as two launches (polar | radial): GPU suite + launch shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s31_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s31_pytest.log | cut -c1-200
timeout 600 python tools/variant_bench.py > gpurun_out/s31_setup_variants.txt 2>&1; cat gpurun_out/s31_setup_variants.txt
