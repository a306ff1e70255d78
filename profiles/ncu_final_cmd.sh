# Round-1 final measurement + profiling recipe (run on the B200 box through gpurun; see /opt/skills/guides/B200_PROFILING.md).
# Every ncu run follows a plain run of the same command line that exited 0.
set -x
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err
python bench.py --steps 5 --warmup 3 --grid explin > gpurun_out/bench_r1_n1_explin.json 2> gpurun_out/bench_r1_n1_explin.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>&1
python profiles/run_configs.py cfg3 cfg4 cfg5 > gpurun_out/configs_r1.jsonl 2> gpurun_out/configs_r1.err; cat gpurun_out/configs_r1.jsonl; tail -3 gpurun_out/configs_r1.err
# launch list of one full-size step (1 chunk stream so that launches are in program order): everything the last step launches
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_all.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo launches_rc=$?
# full counters of the three hot kernels, half-size step (204 pencils, ~2.8 waves)
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1 --zrep 4"
for K in back factor round; do
  SKIP=2; [ $K = round ] && SKIP=3
  $BENCH > gpurun_out/plain_$K.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:bsp_${K}_kernel -s $SKIP -c 1 -o /tmp/prof_${K}_r1 $BENCH > gpurun_out/ncu_$K.log 2>&1
  echo ${K}_rc=$?
  ncu -i /tmp/prof_${K}_r1.ncu-rep --page raw --csv > gpurun_out/prof_${K}_r1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${K}_r1.ncu-rep --page details > gpurun_out/prof_${K}_r1_details.txt 2>/dev/null
  ncu -i /tmp/prof_${K}_r1.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/prof_${K}_r1_source.csv
done
du -sh gpurun_out
