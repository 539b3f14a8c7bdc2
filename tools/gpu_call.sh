timeout 300 python -m pytest tests/test_gpu_uastc.py -m gpu -x -q 2>&1 | tail -2
for lib in libb2bu.so libv_base.so libb2bu.so libv_base.so; do
B2BU_LIBRARY=$PWD/basisu_rs_b200/$lib timeout 200 python bench.py --all-targets --no-cpu-baseline --steps 200 --e2e-steps 2 --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', round(d['ms_per_step']*1e3,1), round(d['ms_per_step_median']*1e3,1), ' '.join('%s=%.0f'%(k,v['us_per_launch']) for k,v in d['extra'].items() if 'shuffled' in k))
"
done
