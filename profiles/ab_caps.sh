# NOTE: the cap_* / stagger / carveout options exist only in commit 33a45dc (work-item launches); results: profiles/ab_caps_r2.txt, DESIGN.md section 12
# A/B of the capped work-item launches (cap_* = block slots per SM) and the staggered chunk streams
# usage: bash profiles/ab_caps.sh > gpurun_out/ab_caps.txt
run() {
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/abc.json 2> gpurun_out/abc.err || { tail -3 gpurun_out/abc.err; return; }
  python - "$*" <<PY
import json,sys
d=json.loads(open("gpurun_out/abc.json").read().strip().splitlines()[-1])
a=d["accuracy"]
print("%-90s %8.0f solves/s %6.2f ms  rounds %s res %.1e orth %.1e info %s"%(sys.argv[1], d["value"], d["ms_per_step"], d["rounds"], a["max_scaled_residual"], a["max_CtSC_minus_I"], a["info_nonzero"]), flush=True)
PY
}
run
run --opt cap_round=2 --opt cap_factor=2 --opt cap_back=2
run --opt cap_round=2 --opt cap_factor=2 --opt cap_back=2 --opt stagger=1 --opt chunk=102
run --opt cap_round=2 --opt cap_factor=1 --opt cap_back=2 --opt stagger=1 --opt chunk=102
run --opt cap_round=2 --opt cap_factor=2 --opt cap_back=2 --opt stagger=1 --opt chunk=68
run --opt cap_round=2 --opt cap_factor=2 --opt cap_back=2 --opt stagger=1 --opt chunk=51 --workers 4
run --opt cap_round=2 --opt stagger=1 --opt chunk=102
run --opt stagger=1 --opt chunk=102
run --opt cap_round=3 --opt cap_factor=1 --opt cap_back=1 --opt stagger=1 --opt chunk=102
run --opt cap_round=1 --opt cap_factor=2 --opt cap_back=3 --opt stagger=1 --opt chunk=102
