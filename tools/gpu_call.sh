timeout 600 python -m pytest tests/test_gpu_etc1s.py -x -q 2>&1 | tail -3
timeout 120 python tools/trace_k2.py libb2bu_k2trace.so 1024 256 8 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 20 --e2e-steps 2 --configs c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(v['entropy_ms'],v['parity_vs_oracle']) for k,v in d['configs']['c4_etc1s'].items() if isinstance(v,dict)})
"
