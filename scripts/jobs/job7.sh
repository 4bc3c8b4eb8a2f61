python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
tail -3 gpurun_out/r2_bench_default_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
print(d['e2e'])
for k, v in d['extra'].items(): print(' ', k, json.dumps(v)[:600])
PY
