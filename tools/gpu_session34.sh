#!/bin/bash
# round 2, GPU session 34 (1 GPU): GPU suite of the final commit (module-level model functions added)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s34_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s34_pytest.log | cut -c1-250
