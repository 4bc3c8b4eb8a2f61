python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['clocks'])
print(d['e2e'])
for k, v in d['extra'].items(): print(' ', k, v.get('value'), v.get('unit'), v.get('roofline_frac'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:mcs_stitch_tiled -s 3 -c 1 -o gpurun_out/prof_r2_final python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
tail -1 gpurun_out/r2_bench_reference_n1.json | cut -c1-400
