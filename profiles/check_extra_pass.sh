# after the "another correction pass" rule: GPU tests, cfg3 / cfg2 / cfg4 bench lines (orthonormality over every pencil, time per step)
python -m pytest tests/test_gpu_solve.py tests/test_gpu_debug.py -m gpu -q --timeout 900 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --config cfg3 --no-cpu-baseline > gpurun_out/bench_r2_cfg3_n1.json 2> gpurun_out/bench_r2_cfg3_n1.err; echo cfg3 rc=$?
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_quick.json 2> gpurun_out/b_quick.err; echo cfg2 rc=$?
python bench.py --steps 3 --warmup 3 --config cfg4 --no-cpu-baseline --no-e2e > gpurun_out/b_cfg4.json 2> gpurun_out/b_cfg4.err; echo cfg4 rc=$?
python - <<PY
import json
for f in ("bench_r2_cfg3_n1","b_quick","b_cfg4"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); a=d["accuracy"]
    print(f, "value %.0f ms %.2f"%(d["value"],d["ms_per_step"]), "res %.1e orth %.1e info %s"%(a["max_scaled_residual"],a["max_CtSC_minus_I"],a["info_nonzero"]), "iters", d["iters"], "sel3", d["selected_third_solve_per_step"], d["kernel_ms_per_step_single_stream"])
PY
