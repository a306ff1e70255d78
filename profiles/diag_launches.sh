BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1"
$BENCH > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 900 --csv --log-file gpurun_out/launches_all.csv $BENCH > gpurun_out/ncu_launches_full.log 2>&1
echo rc=$?
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 600 gpurun_out/bench_r1_n1.json
