timeout 900 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_stitch.py tests/test_stitcher_api.py tests/test_gpu_fuzz.py tests/test_feather.py -m gpu -x -q 2>&1 | tail -5
python scripts/bench_single_call.py
MCS_HOST_THREADS=4 python scripts/bench_single_call.py
MCS_HOST_THREADS=16 python scripts/bench_single_call.py
MCS_HOST_THREADS=1 python scripts/bench_single_call.py
