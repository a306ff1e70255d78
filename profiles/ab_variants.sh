# A/B of compile-time variants built into variants/lib_*.so (BSPATOM_LIB selects the library)
for L in bspatom_b200/libbspatom.so variants/lib_*.so; do
  BSPATOM_LIB=$PWD/$L python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err || { echo "$L FAILED"; tail -3 gpurun_out/ab_tmp.err; continue; }
  python - "$L" <<PY
import json,sys
d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
k=d["kernel_ms_per_step_single_stream"]
print(sys.argv[1], "ms/step %.2f"%d["ms_per_step"], " ".join("%s=%.2f"%(a.replace("bsp_","").replace("_kernel",""),b) for a,b in k.items()), "rounds", d.get("rounds"), "iters", d.get("iters"))
PY
done
