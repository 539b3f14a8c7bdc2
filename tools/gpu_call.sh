for t in astc rgba bc7 etc1 etc2; do bash tools/gpu_profile.sh $t > /dev/null 2>&1; done
bash tools/gpu_profile_etc1s.sh > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 1 --c5-images 32 --c4-slices 16 > gpurun_out/bench_under_ncu.log 2>&1
python tools/trace_pipeline.py astc > gpurun_out/trace_astc.txt 2>&1
python tools/trace_pipeline.py rgba > gpurun_out/trace_rgba.txt 2>&1
python bench.py --all-targets --steps 200 > gpurun_out/bench_all_targets.json 2> gpurun_out/bench_all_targets.err
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_default.csv
