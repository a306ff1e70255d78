# Profiling recipe of round 1 (run on the B200 box through gpurun; see /opt/skills/guides/B200_PROFILING.md).
# Every ncu run follows a plain run of the same command line that exited 0.
set -x
rm -rf gpurun_out/*
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --zrep 1"
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; cat gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo launches_rc=$?
for K in back factor round; do
  $BENCH > gpurun_out/plain_$K.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:bsp_${K}_kernel -s 2 -c 1 -o /tmp/prof_${K}_r1 $BENCH > gpurun_out/ncu_$K.log 2>&1
  echo ${K}_rc=$?
  ncu -i /tmp/prof_${K}_r1.ncu-rep --page raw --csv > gpurun_out/prof_${K}_r1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${K}_r1.ncu-rep --page details > gpurun_out/prof_${K}_r1_details.txt 2>/dev/null
  ncu -i /tmp/prof_${K}_r1.ncu-rep --page source --csv > /tmp/src_$K.csv 2>/dev/null
  # keep the source page only for lines of our own code with stall samples
  head -c 6000000 /tmp/src_$K.csv > gpurun_out/prof_${K}_r1_source.csv
  ls -la /tmp/prof_${K}_r1.ncu-rep
  SZ=$(stat -c %s /tmp/prof_${K}_r1.ncu-rep 2>/dev/null || echo 0)
  if [ "$SZ" -lt 12000000 ]; then cp /tmp/prof_${K}_r1.ncu-rep gpurun_out/; fi
done
du -sh gpurun_out; ls -la gpurun_out/
