timeout 900 python -m pytest tests/test_gpu_uastc.py -m gpu -x -q 2>&1 | tail -2
bash tools/quick_bench.sh
echo "== only mode 8"; python tools/trace_pipeline.py astc libb2bu_trace.so 8 2>&1 | sed -n 1p\;5,11p
