# quick check: GPU tests + one bench line, key numbers only
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 4 --warmup 3 "$@" > gpurun_out/b_quick.json 2> gpurun_out/b_quick.err || tail -5 gpurun_out/b_quick.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b_quick.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["mode"], d["e2e"]["serial"]["ms_each_step"])
print(d["kernel_ms_per_step_single_stream"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["other_kernels"])
print(d["accuracy"], d["rounds"], d["iters"], d["clocks"])
PY
