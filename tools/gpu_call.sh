for lib in libb2bu.so libv_cl2.so libb2bu.so libv_cl2.so; do
B2BU_LIBRARY=$PWD/basisu_rs_b200/$lib timeout 200 python bench.py --no-cpu-baseline --steps 20 --e2e-steps 60 --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'e2e', round(d['e2e']['value'],2), round(d['e2e'].get('frac_of_pcie_ceiling'),3), 'ceiling', round(d['e2e']['pcie_ceiling_gbs'],1), 'parity', d.get('e2e_parity'))
"
done
B2BU_LIBRARY=$PWD/basisu_rs_b200/libv_cl2.so timeout 600 python -m pytest tests/test_gpu_uastc.py -x -q 2>&1 | tail -2
