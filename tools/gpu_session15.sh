#!/bin/bash
# round 2, GPU session 15 (2 GPUs): final kernels: cfg 4 bench with the cfg-5 sweep path forced on, reference arm at N = 2
mkdir -p gpurun_out
KGEO_BENCH_SWEEP=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_cfg4.json 2> gpurun_out/s15_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1200 gpurun_out/s15_bench_2gpu.err | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm_2gpu.json 2> gpurun_out/s15_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_cfg4.json'))
print({k:d[k] for k in ('value','ms_per_step','phases','shard_check','scaling')})
print('roofline', d['roofline']['frac'], d['roofline'].get('aggregate'))
print('e2e', {k:d['e2e'][k] for k in ('value','d2h_gbs','ceiling_gbs','frac_of_ceiling')}, d['e2e']['python']['frac_of_c_abi'])
print('secondary', d['secondary']); print('clocks', d['clocks'])
r=json.load(open('gpurun_out/r02_bench_reference_arm_2gpu.json')); print('ref', r['value'], r['config']['workload'][:60])
PY
