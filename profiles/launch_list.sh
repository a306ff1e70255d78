BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1"
$BENCH > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_all_r2.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo rc=$?
