python tools/trace_pipeline.py astc > gpurun_out/trace_astc_final.txt 2>&1; head -13 gpurun_out/trace_astc_final.txt | cut -c1-300
python tools/trace_pipeline.py rgba > gpurun_out/trace_rgba_final.txt 2>&1
