for f in 0 3 5; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload cfg2_6x1080p --feather $f 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('feather $f', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'launches', d['gpu_launches'], d['parity'], d.get('tiled'))
    elif l.strip(): print(l.rstrip()[:300])
"
done
for f in 0 3; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --feather $f 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8x1080p feather $f', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'launches', d['gpu_launches'], d['parity'], d.get('tiled'))
    elif l.strip(): print(l.rstrip()[:300])
"
done
