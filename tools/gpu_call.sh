set -u
bash tools/gpu_profile.sh astc > /dev/null 2>&1
bash tools/gpu_profile_etc1s.sh > /dev/null 2>&1
timeout 600 python tools/trace_k2.py libb2bu_k2trace.so 1024 1024 64 -1 > gpurun_out/trace_k2_c4.txt 2>&1
timeout 60 ./tools/bin/probe_chase > gpurun_out/probe_chase.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 1 --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
tail -5 gpurun_out/trace_k2_c4.txt
