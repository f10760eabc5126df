#!/bin/sh
# developer tool: build launch-configuration variants of libkgeo_b200.so into kgeo_b200/lib/variants/ (timed on the GPU by
# variant_bench.py).  usage: build_variants.sh "8 2 8 4" ...   (track kernel: warps per CTA, min CTAs per SM, ray groups per CTA; setup min CTAs per SM)
R=$(cd "$(dirname "$0")/../.." && pwd)
mkdir -p "$R/kgeo_b200/lib/variants"
rm -f "$R"/kgeo_b200/lib/variants/*.so
for v in "$@"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=true -diag-suppress 177 -shared -Xcompiler -fPIC \
       -DKG_TRK_WARPS=$1 -DKG_TRK_MINBLOCKS=$2 -DKG_TRK_GROUPS=${3:-8} -DKG_SETUP_MINBLOCKS=${4:-4} -o "$R/kgeo_b200/lib/variants/lib_w$1b$2g${3:-8}s${4:-4}.so" "$R/kgeo_b200/csrc/kgeo_b200.cu" &
done
wait
ls -la "$R/kgeo_b200/lib/variants"
