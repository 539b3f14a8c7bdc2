#!/bin/bash
# Runs bench.py --all-targets over every tuning variant basisu_rs_b200/libv_*.so (plus the default build).
for lib in basisu_rs_b200/libb2bu.so basisu_rs_b200/libv_*.so; do
  name=$(basename $lib .so)
  B2BU_LIBRARY=$PWD/$lib timeout 200 python bench.py --all-targets --no-cpu-baseline --steps 50 --e2e-steps 2 --configs none > gpurun_out/tune_$name.json 2> gpurun_out/tune_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/tune_$name.json"))
    print("$name", " ".join("%s=%.0f" % (k.replace("kat-","").replace("/",":"), v["us_per_launch"]) for k,v in d["extra"].items()))
except Exception as e:
    print("$name FAILED", e)
PY
done
