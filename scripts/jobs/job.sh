( time python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks','parity')})
print(d['roofline'])
print(d['e2e'])
print(d['cpu_baseline'])
for k, v in d['extra'].items(): print(k, v)
PY
