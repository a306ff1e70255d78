python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d1.json 2> gpurun_out/d1.err
python - <<PY
import json
d=json.load(open("gpurun_out/d1.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "wall", d["wall_ms_per_step"])
print(d["kernel_ms_per_step"]); print(d["stage_ms_per_step"]); print(d["gpu_launches"], d["rounds"], d["iters"])
PY
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 333 -c 120 --csv --log-file gpurun_out/launches_full.csv $BENCH > gpurun_out/ncu_launches_full.log 2>&1
echo rc=$?
