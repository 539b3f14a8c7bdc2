export B2BU_LIBRARY=$PWD/basisu_rs_b200/libb2bu_nosort.so
for t in astc bc7 rgba; do
python bench.py --target $t --payload kat-coherent --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$t', d['ms_per_step']*1000,'us')"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none -k regex:uastc_transcode_kernel -s 4 -c 1 python bench.py --target $t --payload kat-coherent --steps 4 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 | grep -E "inst_executed|time_duration|issue_active|no_instruction"
done
