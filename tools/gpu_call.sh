for lib in libb2bu.so libv_p0.so libb2bu.so libv_p0.so; do
B2BU_LIBRARY=$PWD/basisu_rs_b200/$lib timeout 200 python bench.py --all-targets --no-cpu-baseline --steps 200 --e2e-steps 2 --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', round(d['ms_per_step']*1e3,2), round(d['ms_per_step_median']*1e3,2), ' '.join('%s=%.1f'%(k.replace('kat-',''),v['us_per_launch']) for k,v in d['extra'].items() if 'shuffled' in k))
"
done
B2BU_LIBRARY=$PWD/basisu_rs_b200/libv_p0.so timeout 600 python -m pytest tests/test_gpu_uastc.py -x -q 2>&1 | tail -2
