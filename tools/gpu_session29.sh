#!/bin/bash
# round 2, GPU session 29 (1 GPU): final kernels: unmasked parity budget and racecheck refreshed
mkdir -p gpurun_out
timeout 900 python tools/parity_budget.py > gpurun_out/s29_parity_budget.log 2>&1; tail -9 gpurun_out/s29_parity_budget.log | cut -c1-330
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/hostsim/sanitize_case.py > gpurun_out/r02_compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r02_compute_sanitizer_racecheck.log
