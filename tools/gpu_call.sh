set -x
python __graft_entry__.py --smoke 2>&1 | tail -3
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_default.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -4
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","ms_per_step_median","roofline","int_bound","per_path_roofline","e2e","timing","clocks","parity","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
for k,v in d.get("configs",{}).items(): print(k, json.dumps(v)[:1800])
r=json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["ms_per_step"], r["config"]==d["config"])
PY
