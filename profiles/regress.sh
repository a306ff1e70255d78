# regression check of the schedule on the configs that differ most: rounds / iterations / time
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for G in lin explin; do
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --grid $G > gpurun_out/rg.json 2> gpurun_out/rg.err
python - $G <<PY
import json,sys
d=json.loads(open("gpurun_out/rg.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f ms %.2f rounds %s iters %s"%(d["value"], d["ms_per_step"], d["rounds"], d["iters"]), {k: round(v,2) for k,v in d["kernel_ms_per_step_single_stream"].items()})
PY
done
python profiles/run_configs.py cfg3 cfg4 2>&1 | cut -c1-330
