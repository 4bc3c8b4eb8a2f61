export MCS_TILED_FAST=0 MCS_BENCH_ABLATION=1
for order in 0 1; do
for lib in "" $PWD/multicamera_stitching_b200/build/variants/libmcs_nocompute.so; do
for mask in 15 3; do
  echo "order=$order lib=$(basename "$lib") mask=$mask"
  MCS_TILED_ORDER=$order MCS_B200_LIB=$lib MCS_TILED_CLASS_MASK=$mask python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | cut -c1-120
done; done; done
