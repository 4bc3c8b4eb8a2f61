for o in 1 2 4 0; do
MCS_TILED_ORDER=$o timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --workload cfg3_8x2160p --launches-per-step 2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('cfg3 order $o', 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], d['parity'])
"
done
for o in 2 4; do
MCS_TILED_ORDER=$o timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8x1080p order $o', 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], d['parity'])
"
done
