timeout 300 python -m pytest tests/test_feather.py tests/test_gpu_sequence.py tests/test_gpu_stitch.py -m gpu -x -q 2>&1 | tail -4
for sp in 1 0; do
MCS_TILED_BAND_SPREAD=$sp timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --feather 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8x1080p feather 3 spread=$sp', 'ms/step %.4f' % d['ms_per_step'], d['parity'])
"
done
MCS_TILED_BAND_SPREAD=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload cfg2_6x1080p --feather 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('cfg2 feather 3 spread=1', 'ms/step %.4f' % d['ms_per_step'], d['parity'])
"
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8x1080p overwrite', 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'])
"
