# 8-GPU bench lines of the configs (one process per GPU under torchrun); usage: gpurun --gpus 8 -- bash profiles/bench_n8.sh
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_cfg2_n8.json 2> gpurun_out/bench_r2_cfg2_n8.err; echo cfg2 rc=$?
$TR bench.py --gpus 8 --steps 3 --warmup 3 --config cfg3 > gpurun_out/bench_r2_cfg3_n8.json 2> gpurun_out/bench_r2_cfg3_n8.err; echo cfg3 rc=$?
$TR bench.py --gpus 8 --steps 3 --warmup 3 --config cfg4 > gpurun_out/bench_r2_cfg4_n8.json 2> gpurun_out/bench_r2_cfg4_n8.err; echo cfg4 rc=$?
tail -c 600 gpurun_out/bench_r2_cfg2_n8.json; tail -3 gpurun_out/bench_r2_cfg3_n8.err
