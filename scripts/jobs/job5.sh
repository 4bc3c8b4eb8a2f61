for band in 64 1080; do for chunk in 16 32; do
MCS_UPLOAD_BAND=$band timeout 200 python bench.py --no-extra --no-cpu --steps 20 --warmup 5 --chunk $chunk 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']
print('band $band chunk $chunk e2e %.0f h2d MB/step %.0f frac %.3f ceil %.0f' % (e['value'], e['h2d_bytes_per_step']/1e6, e['frac_of_host_ceiling'], e['host_ceiling_panoramas_per_s']))"
done; done
