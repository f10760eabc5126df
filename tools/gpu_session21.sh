#!/bin/bash
# round 2, GPU session 21 (2 GPUs): new public functions + gather variants (copy engine / peer stores / nccl) at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s21_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s21_pytest.log | cut -c1-300
run2() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 6 --warmup 3 2>>gpurun_out/s21_sweep.err | tail -1; }
for g in copy_engine peer_stores; do echo "gather=$g"; KGEO_BENCH_PROFILE=1 KGEO_BENCH_GATHER=$g run2 29601; done > gpurun_out/s21_gather_sweep.txt 2>&1
echo "nccl" >> gpurun_out/s21_gather_sweep.txt; KGEO_BENCH_PROFILE=1 KGEO_B200_GATHER=nccl run2 29711 >> gpurun_out/s21_gather_sweep.txt 2>&1
cut -c1-420 gpurun_out/s21_gather_sweep.txt; grep "rank0\|Error" gpurun_out/s21_sweep.err | tail -5
python - <<'PY'
import time, numpy as np
from kgeo_b200 import equatorial_lensing as el
import contextlib, io, tempfile, torch
spins = np.linspace(0.01, 0.99, 99); incs = np.linspace(1, 89, 89)
I, S = [x.flatten() for x in np.meshgrid(incs, spins)]
req = 1 + np.sqrt(1 - S ** 2)
el.rho_library(S[:10], I[:10] * np.pi / 180, req[:10], 5)
torch.cuda.synchronize(); t0 = time.perf_counter()
vp, rho, st = el.rho_library(S, I * np.pi / 180, req, 5)
dt = time.perf_counter() - t0
print("library: %d configs x 6 orders x 360 angles = %d root finds in %.3f s; configurations with every order converged: %d" % (len(S), rho.size, dt, int((st == 0).all(axis=(1, 2)).sum())))
PY
