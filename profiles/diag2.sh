for V in _v2 _v3; do
echo "== variant $V"
BSPATOM_LIB=$PWD/bspatom_b200/libbspatom$V.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d2.json 2> gpurun_out/d2.err
python - <<PY
import json
d=json.load(open("gpurun_out/d2.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],1))
print("   ", {k:round(v,1) for k,v in d["kernel_ms_per_step_single_stream"].items()}, d["rounds"], d["iters"])
PY
tail -3 gpurun_out/d2.err
done
