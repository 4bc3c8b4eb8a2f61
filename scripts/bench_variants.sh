#!/bin/bash
# Run bench.py (kernel-resident timing only) against every variant library built by build_variant.sh.
cd "$(dirname "$0")/.."
for lib in multicamera_stitching_b200/build/variants/libmcs_*.so; do
  n=$(basename $lib .so)
  MCS_B200_LIB=$PWD/$lib python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$n', 'ms_per_step %.4f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['parity'])
    elif l.strip(): print('$n', l.rstrip())
"
done
