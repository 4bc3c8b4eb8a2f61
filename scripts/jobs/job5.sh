t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
echo "bench wall seconds: $(( $(date +%s) - t0 ))"
tail -2 gpurun_out/r2_bench_default_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
print(d['e2e']['value'], d['e2e']['frac_of_host_ceiling'])
for k, v in d['extra'].items(): print(' ', k, round(v.get('value')), v.get('roofline_frac'))
PY
