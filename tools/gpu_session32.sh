#!/bin/bash
# round 2, GPU session 32 (2 GPUs): the full bench line through the two other delivery paths (NCCL fallback, peer stores)
mkdir -p gpurun_out
for g in nccl peer_stores; do
  if [ $g = nccl ]; then export KGEO_B200_GATHER=nccl; else unset KGEO_B200_GATHER; export KGEO_BENCH_GATHER=peer_stores; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s32_bench_2gpu_$g.json 2> gpurun_out/s32_$g.err; echo "$g rc=$?"
  grep "rank0\|Error" gpurun_out/s32_$g.err | tail -5
  python - <<PY
import json
d=json.load(open('gpurun_out/s32_bench_2gpu_$g.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases','shard_check','gather','gather_check')})
PY
done
