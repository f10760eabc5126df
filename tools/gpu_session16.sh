#!/bin/bash
# round 2, GPU session 16 (8 GPUs): BASELINE cfg 4 as itself on the whole box + the cfg-5 sweep; dist_check
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s16_topo.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu_cfg4.json 2> gpurun_out/s16_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 1500 gpurun_out/s16_bench_8gpu.err | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py 2>&1 | tail -2
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_8gpu_cfg4.json'))
print({k:d[k] for k in ('value','ms_per_step','phases','shard_check','scaling')})
print('roofline', d['roofline']['frac'], d['roofline'].get('aggregate'))
print('e2e', {k:d['e2e'][k] for k in ('value','d2h_gbs','ceiling_gbs','frac_of_ceiling','host_cpu_affinity')}, d['e2e']['python']['frac_of_c_abi'])
print('secondary', d['secondary']); print('clocks', d['clocks'])
PY
