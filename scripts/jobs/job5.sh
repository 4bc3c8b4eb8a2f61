timeout 500 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k blend 2>&1 | tail -15
