timeout 600 python -m pytest tests/test_feather.py tests/test_gpu_sequence.py -m gpu -x -q 2>&1 | tail -3
for f in 0 3; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --feather $f 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8x1080p feather $f', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'launches', d['gpu_launches'], d['parity'])
    elif l.strip(): print(l.rstrip()[:300])
"
done
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload cfg2_6x1080p --feather 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('cfg2 feather 3', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'launches', d['gpu_launches'], d['parity'])
    elif l.strip(): print(l.rstrip()[:300])
"
