#!/bin/bash
# round 2, GPU session 8 (2 GPUs): full GPU tests incl. the multi-GPU one, setup barrier variants, dist_check, bench.py --gpus 2 (cfg 4)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s8_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log
tail -6 gpurun_out/s8_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/variant_bench.py > gpurun_out/s8_variants.txt 2>&1
cat gpurun_out/s8_variants.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s8_bench_2gpu.json 2> gpurun_out/s8_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 2500 gpurun_out/s8_bench_2gpu.err | tail -12
head -c 7000 gpurun_out/s8_bench_2gpu.json
