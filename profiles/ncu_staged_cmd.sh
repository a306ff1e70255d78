# ncu --set full of the three hot kernels after the shared-memory staging change (half-size step, 204 pencils).
# Every ncu run follows a plain run of the same command line that exited 0.
set -x
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1 --zrep 4"
for K in round factor back; do
  SKIP=2; [ $K = round ] && SKIP=7
  $BENCH > gpurun_out/plain_$K.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:bsp_${K}_kernel -s $SKIP -c 1 -o /tmp/prof_${K}_r1b $BENCH > gpurun_out/ncu_$K.log 2>&1
  echo ${K}_rc=$?
  ncu -i /tmp/prof_${K}_r1b.ncu-rep --page raw --csv > gpurun_out/prof_${K}_r1b_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${K}_r1b.ncu-rep --page details > gpurun_out/prof_${K}_r1b_details.txt 2>/dev/null
  ncu -i /tmp/prof_${K}_r1b.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/prof_${K}_r1b_source.csv
done
du -sh gpurun_out
