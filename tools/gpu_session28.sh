#!/bin/bash
# round 2, GPU session 28 (1 GPU): launch shapes of the emission kernel
mkdir -p gpurun_out
timeout 600 python tools/image_bench.py > gpurun_out/s28_image_variants.txt 2>&1; cat gpurun_out/s28_image_variants.txt
