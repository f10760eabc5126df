#!/bin/bash
# round 2, GPU session 18 (2 GPUs): image kernel with the gather fused in (peer stores over NVLink): tests, block-count sweep, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s18_pytest.log
timeout 300 python tools/image_bench.py > gpurun_out/s18_image_bench.txt 2>&1; tail -5 gpurun_out/s18_image_bench.txt
run2() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 6 --warmup 3 2>>gpurun_out/s18_sweep.err | tail -1; }
for nb in 0 148 296 592 1184; do echo "img_blocks=$nb"; KGEO_BENCH_PROFILE=1 KGEO_BENCH_IMG_BLOCKS=$nb run2 $((29600+nb%97)); done > gpurun_out/s18_block_sweep.txt 2>&1
echo "nccl" >> gpurun_out/s18_block_sweep.txt; KGEO_BENCH_PROFILE=1 KGEO_B200_GATHER=nccl run2 29711 >> gpurun_out/s18_block_sweep.txt 2>&1
cat gpurun_out/s18_block_sweep.txt | cut -c1-400; tail -5 gpurun_out/s18_sweep.err
KGEO_BENCH_SWEEP=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s18_bench_2gpu.json 2> gpurun_out/s18_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/s18_bench_2gpu.err | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/s18_bench_2gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases','shard_check','gather','gather_check')})
print('roofline', d['roofline']['frac'], d['roofline'].get('aggregate'))
print('secondary', d['secondary'])
PY
