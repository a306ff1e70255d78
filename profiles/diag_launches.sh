BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1"
$BENCH > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 36 --csv --log-file gpurun_out/launches_full.csv $BENCH > gpurun_out/ncu_launches_full.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_launches_full.log | cut -c1-300
