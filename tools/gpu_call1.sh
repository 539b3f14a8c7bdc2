set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
tools/bin/probe_pipes > gpurun_out/probe_pipes_r02.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_a.log 2>&1; tail -3 gpurun_out/pytest_r02_a.log
bash tools/quick_bench.sh
python tools/trace_pipeline.py astc > gpurun_out/trace_astc_r02_a.txt 2>&1
python tools/trace_pipeline.py rgba > gpurun_out/trace_rgba_r02_a.txt 2>&1
cat gpurun_out/probe_pipes_r02.txt
