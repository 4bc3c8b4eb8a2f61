python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')})
print(d['e2e'])
for k, v in d['extra'].items(): print(k, v.get('value'), v.get('unit'))
PY
