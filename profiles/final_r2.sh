# final check of the round-2 head: GPU tests, smoke, bench line (refreshes bench_r2_n1.json with roofline.traffic from roofline_r2.json)
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_r2.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -2 gpurun_out/smoke_r2.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -2 gpurun_out/bench_r2_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["pcie"]["frac_of_pcie_floor"], "sel", d["e2e"]["selected"]["value"], "traffic", d["roofline"]["traffic"], d["clocks"])
PY
