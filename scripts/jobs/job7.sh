timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
timeout 600 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
tail -3 gpurun_out/r2_bench_default_n1.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/r2_bench_reference_n1.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['cpu_baseline']['cores'])
d = json.loads(open('gpurun_out/r2_bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
print(d['e2e']); print(d['cpu_baseline'])
for k, v in d['extra'].items(): print(' ', k, json.dumps(v)[:420])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mcs_stitch_tiled -s 3 -c 1 -o gpurun_out/prof_r2_final python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log | cut -c1-100
timeout 300 python bench.py --feather 3 --no-extra --no-cpu --steps 10 --warmup 3 > gpurun_out/r2_bench_feather3.json 2>/dev/null
MCS_TILED_BAND=0 timeout 300 python bench.py --feather 3 --no-extra --no-cpu --steps 10 --warmup 3 > gpurun_out/r2_bench_feather3_twopass.json 2>/dev/null
python - <<'PY'
import json
for f in ('r2_bench_feather3','r2_bench_feather3_twopass'):
    d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    print(f, round(d['value']), d['ms_per_step'], d['gpu_launches'], d['parity'], 'e2e', round(d['e2e']['value']), d['e2e']['h2d_bytes_per_step'])
PY
