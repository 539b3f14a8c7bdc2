timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --all-targets --no-cpu-baseline --steps 200 --e2e-steps 2 --configs none 2>/dev/null | tail -1 > gpurun_out/all_targets.json
python -c "
import json
d=json.loads(open('gpurun_out/all_targets.json').read())
print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_median']*1e3,2), ' '.join('%s=%.1f'%(k.replace('kat-',''),v['us_per_launch']) for k,v in d['extra'].items()))
"
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_default.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_default.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])
"
