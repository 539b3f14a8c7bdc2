python bench.py --gpus 1 --steps 20 --warmup 5 --configs none > gpurun_out/b1.json 2>/dev/null
python bench.py --gpus 1 --steps 200 --warmup 5 --configs none --all-targets > gpurun_out/b2.json 2>/dev/null
python - <<'PY'
import json
for f in ("b1","b2"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, "ms", round(d["ms_per_step"]*1e3,2), "median(separated)", round(d["ms_per_step_median"]*1e3,2), d["timing"]["per_step_us_min"], d["timing"]["per_step_us_max"], "frac", round(d["roofline"]["frac"],3), round(d["per_path_roofline"]["frac_of_per_path_roofline"],3), {k.split("/")[0]:round(v["us_per_launch"],1) for k,v in d.get("extra",{}).items() if "shuffled" in k})
PY
