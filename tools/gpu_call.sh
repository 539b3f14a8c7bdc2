for lib in libb2bu.so libv_narrow.so libb2bu.so libv_narrow.so; do
B2BU_LIBRARY=$PWD/basisu_rs_b200/$lib timeout 300 python bench.py --no-cpu-baseline --steps 20 --e2e-steps 2 --configs c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', {k:(round(v['entropy_ms'],2),v['parity_vs_oracle']) for k,v in d['configs']['c4_etc1s'].items() if isinstance(v,dict)})
"
done
