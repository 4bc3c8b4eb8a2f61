timeout 200 python -m pytest tests/test_cabi_errors.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python scripts/probe_wc_upload.py 2>&1 | tail -10
