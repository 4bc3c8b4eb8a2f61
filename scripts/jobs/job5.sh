for b in 96 128; do
MCS_TILED_FRAME_BLOCK=$b timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --batch $b 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('batch = frame block = $b', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'frac %.4f' % d['roofline']['frac'], d['parity'])
"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mcs_stitch_tiled -s 3 -c 1 -o gpurun_out/prof_r2_final python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log | cut -c1-100
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
print(d['e2e']['value'], d['e2e']['frac_of_host_ceiling'])
PY
