#!/bin/bash
# round 2, GPU session 30 (4 GPUs): final cfg-4 line at N = 4 (roofline from the re-captured summary) 
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_4gpu_cfg4.json 2> gpurun_out/s30_bench_4gpu.err; echo "bench4 rc=$?"
grep -v "^$" gpurun_out/s30_bench_4gpu.err | grep "rank0\|Error" | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_4gpu_cfg4.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases','shard_check','gather','gather_check')})
print('roofline', d['roofline']['frac'], d['roofline'].get('aggregate'))
print('secondary', d['secondary'])
print('e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'], d['e2e']['python']['frac_of_c_abi'])
PY
