#!/bin/bash
# Run under gpurun: ncu capture of the ETC1S entropy-decode kernel (K2) on a 256x256-block x 16-slice payload
ncu --set full --clock-control none --import-source on -k regex:etc1s_entropy -s 1 -c 1 -f -o gpurun_out/prof_etc1s \
    python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 1 --configs c4 --c4-blocks 256 --c4-slices 16 > gpurun_out/prof_etc1s.log 2>&1
tail -3 gpurun_out/prof_etc1s.log
