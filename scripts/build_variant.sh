#!/bin/bash
# Build an experimental variant of libmcs_b200.so:  scripts/build_variant.sh NAME -DTILED_MIN_CTAS=3 ...
# -> multicamera_stitching_b200/build/variants/libmcs_NAME.so   (select it with MCS_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=multicamera_stitching_b200/build/variants
mkdir -p $out/$name
for f in mcs_plan mcs_tiles mcs_stitch mcs_stitch_tiled mcs_match mcs_ransac mcs_resize mcs_hostio mcs_refit; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false \
       -Xcompiler -fPIC,-O2,-ffp-contract=off -Xptxas -v -I include -I multicamera_stitching_b200/csrc "$@" \
       -c multicamera_stitching_b200/csrc/$f.cu -o $out/$name/$f.o > $out/$name/$f.log 2>&1 &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libmcs_$name.so $out/$name/*.o
grep -A2 "stitch_tiled_kernelILi3" $out/$name/mcs_stitch_tiled.log | grep -E "spill|Used"
