#!/bin/bash
# round 2, GPU session 33 (1 GPU): GPU suite of the final commit (n_angular_turnings added) + smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s33_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s33_pytest.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
