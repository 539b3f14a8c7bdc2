# usage: tools/gpu_call_n.sh N  -- the driver's multi-rank launch of bench.py
N=$1
timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], d["ms_per_step_median"])
print("e2e", json.dumps(d["e2e"])[:900])
c5=d.get("configs",{}).get("c5_mixed_batch"); print("c5", json.dumps(c5)[:1500])
PY
