# direct-store experiment against the base library
for lib in "" multicamera_stitching_b200/build/variants/libmcs_direct.so; do
  echo "lib=$lib"
  MCS_B200_LIB=${lib:+$PWD/$lib} python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms/step %.4f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['parity'])
    elif l.strip(): print(l.rstrip()[:300])
"
done
