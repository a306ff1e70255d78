for V in "" _v2 _v3; do
  echo "== variant '$V'"
  BSPATOM_LIB=$PWD/bspatom_b200/libbspatom$V.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab$V.json 2> gpurun_out/ab$V.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab$V.json"))
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["e2e"]["ms_per_step"], d["e2e"]["wall_ms_last_step"])
print(d["kernel_ms_per_step"], "rounds", d["rounds"], "iters", d["iters"])
PY
  tail -2 gpurun_out/ab$V.err
done
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
