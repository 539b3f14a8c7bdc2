timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?
timeout 600 python bench.py --all-targets --no-cpu-baseline --steps 200 --e2e-steps 20 --configs none > gpurun_out/bench_all_targets.json 2> gpurun_out/bench_all_targets.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("default:", d["value"], d["ms_per_step"], d["ms_per_step_median"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_pcie_ceiling"), "frac", d["roofline"]["frac"], d["per_path_roofline"]["frac_of_per_path_roofline"], "cpu", d["cpu_baseline"]["value"], "launches", d.get("gpu_launches"))
c=d["configs"]
print("c3", c["c3_bc7_mip_chain"]["us_per_chain"], "c4", {k:(round(v['entropy_ms'],2), v['parity_vs_oracle']) for k,v in c['c4_etc1s'].items() if isinstance(v,dict)})
a=json.loads(open("gpurun_out/bench_all_targets.json").read().strip().splitlines()[-1]); print("all:", {k:round(v['us_per_launch'],1) for k,v in a['extra'].items() if 'shuffled' in k}, a['ms_per_step'])
PY
