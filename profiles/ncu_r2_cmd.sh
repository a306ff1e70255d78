# Round-2 measurement + profiling recipe (run on the B200 box through gpurun; see /opt/skills/guides/B200_PROFILING.md).
# Every ncu run follows a plain run of the same command line that exited 0.  Results are folded into profiles/ by
# `python profiles/collect.py r2`.
set -x
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_r2.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -2 gpurun_out/smoke_r2.log
# north star stages (ii)+(iii) as a stand-alone A/B: band -> tridiagonal bulge chasing and one tridiagonal Sturm round, 408 matrices
[ -x profiles/micro/tridiag_ab ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o profiles/micro/tridiag_ab profiles/micro/tridiag_ab.cu
timeout 300 ./profiles/micro/tridiag_ab 1000 6 408 3 gpurun_out/tri.bin > gpurun_out/tridiag_ab_r2.json 2> gpurun_out/tridiag_ab_r2.err; cat gpurun_out/tridiag_ab_r2.json
timeout 300 python profiles/micro/tridiag_ab_check.py gpurun_out/tri.bin > gpurun_out/tridiag_ab_check_r2.json 2>&1; cat gpurun_out/tridiag_ab_check_r2.json; rm -f gpurun_out/tri.bin
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -2 gpurun_out/bench_r2_n1.err
# launch list of one full-size step (1 chunk stream so that launches are in program order): everything the last step launches
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_all_r2.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo launches_rc=$?
# full counters of ONE full-width launch of the three hot kernels, half-size step (204 pencils, ~2.8 waves).  Per step the
# schedule launches 5 factor / 5 back kernels (2 full-width, 1 compacted, 2 no-ops) and 26 round kernels.
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workers 1 --zrep 4"
for K in back factor round; do
  SKIP=5; [ $K = factor ] && SKIP=6; [ $K = round ] && SKIP=29
  $BENCH > gpurun_out/plain_$K.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:bsp_${K}_kernel -s $SKIP -c 1 -o /tmp/prof_${K}_r2 $BENCH > gpurun_out/ncu_$K.log 2>&1
  echo ${K}_rc=$?
  ncu -i /tmp/prof_${K}_r2.ncu-rep --page raw --csv > gpurun_out/prof_${K}_r2_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${K}_r2.ncu-rep --page details > gpurun_out/prof_${K}_r2_details.txt 2>/dev/null
done
du -sh gpurun_out
