for l in libb2bu.so libb2bu_c17s4.so libb2bu_c17s6.so libb2bu_c18s4.so libb2bu_c20s3.so; do
B2BU_LIBRARY=$PWD/basisu_rs_b200/$l python bench.py --no-cpu-baseline --configs none --steps 20 --e2e-steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$l', round(d['e2e']['value'],1), d['e2e']['launches'])"
done
