timeout 600 python -m pytest tests/test_recorded.py tests/test_gpu_sequence.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, json
sys.argv=['bench.py']
import bench, torch
torch.cuda.set_device(0)
print(json.dumps(bench.extra_recorded(torch.device('cuda',0))))
PY
