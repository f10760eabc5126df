#!/bin/bash
# round 2, GPU session 24 (2 GPUs): GPU suite (multi-GPU test included) + final cfg-4 line at N = 2 with the roofline of the re-captured summary
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s24_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s24_pytest.log | cut -c1-250
KGEO_BENCH_SWEEP=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_cfg4.json 2> gpurun_out/s24_bench_2gpu.err; echo "bench2 rc=$?"
grep -v "^$" gpurun_out/s24_bench_2gpu.err | grep "rank0\|Error" | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_cfg4.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases','shard_check','gather','gather_check')})
print('roofline', d['roofline']['frac'], d['roofline'].get('aggregate'))
print('secondary', d['secondary'])
print('e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'], d['e2e']['python']['frac_of_c_abi'])
PY
