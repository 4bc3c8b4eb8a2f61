timeout 300 python -m pytest tests/test_gpu_stitch.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for fast in 0 32; do echo "fastmax $fast"; MCS_TILED_FAST_MAX=$fast python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | cut -c1-330; done
export MCS_TILED_FAST_MAX=0
for fb in 32 16; do echo "fb $fb"; MCS_TILED_FRAME_BLOCK=$fb python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | cut -c1-130; done
for w in ns_8x1080p cfg1_3x720p cfg3_8x2160p; do echo $w; python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | cut -c1-200; done
