#!/bin/bash
# round 2, GPU session 19 (2 GPUs): fused image + peer-store gather: dist_check, image kernel timing, full cfg-4 bench line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|image [0-9]" | tail -5
timeout 300 python tools/image_bench.py > gpurun_out/s19_image_bench.txt 2>&1; tail -2 gpurun_out/s19_image_bench.txt
KGEO_BENCH_SWEEP=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s19_bench_2gpu.json 2> gpurun_out/s19_bench_2gpu.err; echo "bench2 rc=$?"
grep -v "^$" gpurun_out/s19_bench_2gpu.err | grep "rank0\|Error" | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/s19_bench_2gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases','shard_check','gather','gather_check')})
print('roofline', d['roofline']['frac'], d['roofline'].get('aggregate'))
print('secondary', d['secondary'])
print('e2e', d['e2e']['value'], d['e2e']['python']['frac_of_c_abi'])
PY
