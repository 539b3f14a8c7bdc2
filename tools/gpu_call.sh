timeout 300 python -m pytest tests/test_gpu_uastc.py -m gpu -x -q 2>&1 | tail -2
timeout 600 bash tools/tune_variants.sh 2>&1 | sed -E 's/(rgba|astc|bc7|etc1|etc2):(random)=[0-9]+ //g'
