#!/bin/bash
# round 2, GPU session 4: ring-buffer TMA variants; the new bench.py (N = 1) end to end; the reference arm
mkdir -p gpurun_out
timeout 600 python tools/variant_bench.py > gpurun_out/s4_variants.txt 2>&1
cat gpurun_out/s4_variants.txt
timeout 300 python -m pytest tests/test_gpu_tracks.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/s4_bench_1gpu.err | tail -15
head -c 6000 gpurun_out/s4_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s4_bench_ref.json 2> gpurun_out/s4_bench_ref.err; echo "ref rc=$?"
head -c 1500 gpurun_out/s4_bench_ref.json
