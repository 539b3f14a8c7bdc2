#!/bin/bash
# Run under gpurun: a light ncu pass (a dozen counters, not --set full) over the pipelined kernel of each listed target.
# usage: tools/gpu_metrics.sh astc bc7 ...   -> gpurun_out/metrics_<target>.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_read.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
for T in "$@"; do
  ncu --metrics $M --clock-control none -k regex:uastc_sorted -s 4 -c 1 --csv --log-file gpurun_out/metrics_$T.csv \
      python bench.py --target $T --steps 4 --warmup 3 --e2e-steps 1 --no-cpu-baseline --configs none > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/metrics_$T.csv")) if len(r)>10]
h=rows[0]; i=h.index("Metric Name"); v=h.index("Metric Value")
print("$T:", "; ".join("%s=%s"%(r[i].split("__")[-1][:48], r[v]) for r in rows[1:]))
PY
done
