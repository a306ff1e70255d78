python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
for F in "--workers 1" "--workers 2" "--workers 3" "--workers 4"; do
python bench.py --steps 6 --warmup 3 --no-cpu-baseline $F > gpurun_out/d2.json 2> gpurun_out/d2.err
python - <<PY
import json
d=json.load(open("gpurun_out/d2.json"))
print("flags '$F' value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), d["e2e"]["ms_each_step"], d["e2e"]["wall_ms_last_step"]), "launches", d["gpu_launches"], d["rounds"], d["iters"])
PY
tail -3 gpurun_out/d2.err
done
