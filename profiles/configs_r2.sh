# N=1 bench lines of the other configs and the reference arm, final round-2 code
python bench.py --steps 3 --warmup 3 --config cfg3 > gpurun_out/bench_r2_cfg3_n1.json 2> gpurun_out/bench_r2_cfg3_n1.err; echo cfg3 rc=$?
python bench.py --steps 3 --warmup 3 --config cfg4 --no-cpu-baseline > gpurun_out/bench_r2_cfg4_n1.json 2> gpurun_out/bench_r2_cfg4_n1.err; echo cfg4 rc=$?
python bench.py --steps 5 --warmup 3 --config cfg5 > gpurun_out/bench_r2_cfg5_n1.json 2> gpurun_out/bench_r2_cfg5_n1.err; echo cfg5 rc=$?
python bench.py --steps 5 --warmup 3 --grid explin --no-cpu-baseline > gpurun_out/bench_r2_n1_explin.json 2> gpurun_out/bench_r2_n1_explin.err; echo explin rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>&1; echo ref rc=$?
for f in cfg3_n1 cfg4_n1 cfg5_n1 n1_explin; do tail -c 300 gpurun_out/bench_r2_$f.json; echo; done
