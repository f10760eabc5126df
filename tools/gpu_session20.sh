#!/bin/bash
# round 2, GPU session 20 (8 GPUs): fused image + peer-store gather at N = 8: block-cap sweep (kernel-only mode), NCCL variant, full line
mkdir -p gpurun_out
run8() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 2>>gpurun_out/s20_sweep.err | tail -1; }
for nb in 0 1184; do echo "img_blocks=$nb"; KGEO_BENCH_PROFILE=1 KGEO_BENCH_IMG_BLOCKS=$nb run8 $((29600+nb%97)); done > gpurun_out/s20_block_sweep.txt 2>&1
echo "nccl" >> gpurun_out/s20_block_sweep.txt; KGEO_BENCH_PROFILE=1 KGEO_B200_GATHER=nccl run8 29711 >> gpurun_out/s20_block_sweep.txt 2>&1
cut -c1-420 gpurun_out/s20_block_sweep.txt; grep "rank0\|Error" gpurun_out/s20_sweep.err | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s20_bench_8gpu.json 2> gpurun_out/s20_bench_8gpu.err; echo "bench8 rc=$?"
grep -v "^$" gpurun_out/s20_bench_8gpu.err | grep "rank0\|Error" | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/s20_bench_8gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases','shard_check','gather','gather_check')})
print('secondary', d['secondary'])
print('e2e', d['e2e']['value'], d['e2e']['python']['frac_of_c_abi'])
PY
