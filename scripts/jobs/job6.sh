timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --feather 3 --no-extra --steps 10 --warmup 3 > gpurun_out/r2_bench_feather3.json 2> gpurun_out/r2_bench_feather3.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_feather3.json').read().strip().splitlines()[-1])
print('feather3', {k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['parity'])
print(d['e2e'])
PY
MCS_TILED_BAND=0 python bench.py --feather 3 --no-extra --no-cpu --steps 10 --warmup 3 > gpurun_out/r2_bench_feather3_twopass.json 2> gpurun_out/r2_bench_feather3_twopass.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_feather3_twopass.json').read().strip().splitlines()[-1])
print('feather3 two-pass', {k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['parity'])
print(d['e2e'])
PY
