#!/bin/sh
# developer tool: build launch-configuration variants of libkgeo_b200.so into kgeo_b200/lib/variants/ (timed on the GPU by
# variant_bench.py).  usage: build_variants.sh "name -DKG_TRK_P1_WARPS=4 -DKG_TRK_P1_MINBLOCKS=8 ..." ...
# knobs: KG_TRK_WARPS / KG_TRK_MINBLOCKS (fused + fix-up passes), KG_TRK_HOT_MINBLOCKS, KG_TRK_P1_* (polar pass),
#        KG_TRK_P2_* (radial pass), KG_TRK_GROUPS (4-ray groups per CTA), KG_SETUP_MINBLOCKS
R=$(cd "$(dirname "$0")/../.." && pwd)
mkdir -p "$R/kgeo_b200/lib/variants"
rm -f "$R"/kgeo_b200/lib/variants/*.so "$R"/kgeo_b200/lib/variants/*.log
for v in "$@"; do
  set -- $v
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=true -diag-suppress 177 -shared -Xcompiler -fPIC \
       -Xptxas -v "$@" -o "$R/kgeo_b200/lib/variants/lib_$name.so" "$R/kgeo_b200/csrc/kgeo_b200.cu" > "$R/kgeo_b200/lib/variants/lib_$name.log" 2>&1 &
done
wait
for f in "$R"/kgeo_b200/lib/variants/*.log; do
  echo "== $f"; grep -A3 "Compiling entry function '_Z12track_kernelILb1" "$f" | grep "Compiling\|spill\|Used" | sed -e "s/for .sm_100a.//" -e "s/.*track_kernelILb1ELi\(.\).*/pass \1/"
done
