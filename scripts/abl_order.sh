export MCS_TILED_FAST=0 MCS_BENCH_ABLATION=1
for order in 0 1 2 4 8 16; do
for lib in "" $PWD/multicamera_stitching_b200/build/variants/libmcs_nocompute.so; do
  echo "order=$order lib=$(basename "$lib")"
  MCS_TILED_ORDER=$order MCS_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | cut -c1-120
done
MCS_TILED_ORDER=$order ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:mcs_stitch_tiled -s 3 -c 1 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | grep -E "dram__bytes|gpu__time"
done
