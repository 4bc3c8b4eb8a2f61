timeout 400 python -m pytest tests/test_feather.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8x1080p overwrite', 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'])
"
done
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --feather 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8x1080p feather 3', 'ms/step %.4f' % d['ms_per_step'], d['parity'])
"
