n=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
tail -2 gpurun_out/r2_bench_n$n.err
python - $n <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open('gpurun_out/r2_bench_n%s.json' % n).read().strip().splitlines()[-1])
print(n, {k: d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'])
print(d['e2e'])
for k, v in d['extra'].items(): print(' ', k, v.get('value'), v.get('unit'), v.get('e2e_panoramas_per_s'))
PY
