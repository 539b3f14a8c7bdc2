#!/bin/bash
# Run under gpurun: tools/gpu_profile.sh <target> [payload]  -> gpurun_out/launches_<target>.csv, gpurun_out/prof_<target>.ncu-rep
set -u
T=${1:-astc}; P=${2:-kat-shuffled}
CMD="python bench.py --target $T --payload $P --steps 4 --warmup 3 --e2e-steps 1 --no-cpu-baseline --configs none"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$T.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:uastc_sorted -s 4 -c 1 -f -o gpurun_out/prof_$T $CMD > gpurun_out/prof_$T.log 2>&1
tail -3 gpurun_out/prof_$T.log
