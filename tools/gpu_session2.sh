#!/bin/bash
# round 2, GPU session 2: hot-kernel variants (register caps, pair loads), ncu source-level profile of the hot kernel,
# ncu of the image / lensing kernels exported as CSV on the box (the reports themselves are too large to bring back)
mkdir -p gpurun_out
timeout 600 python tools/variant_bench.py > gpurun_out/s2_variants.txt 2>&1
cat gpurun_out/s2_variants.txt
for v in b_pairs_w8_b4 d_ns1_w4_b9; do
  KGEO_B200_LIB=$PWD/kgeo_b200/lib/variants/lib_$v.so timeout 300 python -m pytest tests/test_gpu_tracks.py -q -k "two_launch or against_reference or oracle_seeded" 2>&1 | tail -2 | sed "s/^/$v: /"
done
KGEO_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_hot_kernel -s 3 -c 1 -o gpurun_out/r02_hot_ns1 python bench.py --steps 1 --warmup 3 > gpurun_out/s2_ncu_hot.log 2>&1
PROFILE_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"image_kernel|r_equatorial_kernel" -o /tmp/r02_image_base python tools/profile_image.py 1 > gpurun_out/s2_ncu_image.log 2>&1
ncu -i /tmp/r02_image_base.ncu-rep --page raw --csv > gpurun_out/r02_image_base_raw.csv 2>/dev/null
ncu -i /tmp/r02_image_base.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_image_base_source.csv 2>/dev/null
gzip -9 gpurun_out/r02_image_base_source.csv
ls -la gpurun_out /tmp/r02_image_base.ncu-rep
du -sh gpurun_out
