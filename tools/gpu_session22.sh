#!/bin/bash
# round 2, GPU session 22 (8 GPUs): gather variants at N = 8 (kernel-only mode), full line with the faster one, dist_check at 8 ranks
mkdir -p gpurun_out
run8() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 2>>gpurun_out/s22_sweep.err | tail -1; }
for g in copy_engine peer_stores; do echo "gather=$g"; KGEO_BENCH_PROFILE=1 KGEO_BENCH_GATHER=$g run8 29601; done > gpurun_out/s22_gather_sweep.txt 2>&1
cut -c1-420 gpurun_out/s22_gather_sweep.txt; grep "rank0\|Error" gpurun_out/s22_sweep.err | tail -5
best=$(python - <<'PY'
import json
lines=[l for l in open('gpurun_out/s22_gather_sweep.txt') if l.startswith('{')]
d=[json.loads(l) for l in lines]
print(min(d, key=lambda x: x['ms_per_step'])['gather'] if d else 'copy_engine')
PY
)
echo "best=$best"
KGEO_BENCH_GATHER=$best timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s22_bench_8gpu.json 2> gpurun_out/s22_bench_8gpu.err; echo "bench8 rc=$?"
grep -v "^$" gpurun_out/s22_bench_8gpu.err | grep "rank0\|Error" | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py 2>&1 | grep "dist_check\|Error\|assert" | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/s22_bench_8gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases','shard_check','gather','gather_check')})
print('secondary', d['secondary'])
print('e2e', d['e2e']['value'], d['e2e']['python']['frac_of_c_abi'])
PY
