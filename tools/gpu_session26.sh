#!/bin/bash
# round 2, GPU session 26 (1 GPU): experiments: harmonic cut-off and image launch-shape variants; source-level profile of the image kernel
mkdir -p gpurun_out
timeout 600 python tools/variant_bench.py > gpurun_out/s26_hot_variants.txt 2>&1; cat gpurun_out/s26_hot_variants.txt
timeout 600 python tools/image_bench.py > gpurun_out/s26_image_variants.txt 2>&1; cat gpurun_out/s26_image_variants.txt
for v in cut17 cut16; do echo "== parity $v"; KGEO_B200_LIB=$PWD/kgeo_b200/lib/variants/lib_$v.so timeout 600 python -m pytest tests/test_gpu_tracks.py tests/test_gpu_images.py -m gpu -q 2>&1 | tail -4 | cut -c1-200; done > gpurun_out/s26_parity.txt 2>&1; cat gpurun_out/s26_parity.txt
PROFILE_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"image_kernel" -s 2 -c 1 -o /tmp/img_src python tools/profile_image.py 1 > gpurun_out/s26_ncu_img.log 2>&1
ncu -i /tmp/img_src.ncu-rep --page source --csv --print-source cuda > gpurun_out/s26_image_source_cuda.csv 2>/dev/null
ncu -i /tmp/img_src.ncu-rep --page source --csv --print-source sass > gpurun_out/s26_image_source_sass.csv 2>/dev/null
ls -la gpurun_out/s26_*
