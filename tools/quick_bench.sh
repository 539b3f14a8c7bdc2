#!/bin/bash
# quick device-resident timing of all targets (kat-shuffled); usage: tools/quick_bench.sh [lib.so]
[ -n "$1" ] && export B2BU_LIBRARY=$PWD/basisu_rs_b200/$1
python bench.py --all-targets --no-cpu-baseline --configs none --steps 50 --e2e-steps 5 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err || tail -5 gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json"))
print("astc main us", round(d["ms_per_step"] * 1e3, 1), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1), "parity", d["parity"])
print({k: round(v["us_per_launch"], 1) for k, v in d["extra"].items() if "shuffled" in k})
PY
