python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
for F in "--workers 2" "--workers 1"; do
python bench.py --steps 4 --warmup 3 --no-cpu-baseline $F > gpurun_out/d2.json 2> gpurun_out/d2.err
python - <<PY
import json
d=json.load(open("gpurun_out/d2.json"))
print("flags '$F' value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "e2e", d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), d["e2e"]["ms_each_step"]))
print("   ", {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d["rounds"], d["iters"], d["gpu_launches"])
PY
tail -3 gpurun_out/d2.err
done
