python -m pytest tests/test_gpu_solve.py -m gpu -q --timeout 900 2>&1 | tail -2
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/d2.json 2> gpurun_out/d2.err
python - <<PY
import json
d=json.load(open("gpurun_out/d2.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), d["e2e"]["ms_each_step"], d["e2e"]["wall_ms_last_step"]))
PY
tail -3 gpurun_out/d2.err
