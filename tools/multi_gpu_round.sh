#!/bin/bash
# Runs on an 8-GPU box (gpurun --gpus 8): scaling lines of every workload, the 2-rank CLI test, the Tecnick schedule sweep.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PORT=29600
run() {  # run <tag> <nproc> <bench args>
  PORT=$((PORT+1))
  timeout 400 $TR --nproc-per-node $2 --master-port $PORT bench.py --gpus $2 $3 --no-cpu > gpurun_out/bench_r2_$1.log 2>&1
  grep '^{' gpurun_out/bench_r2_$1.log | tail -1 | cut -c1-160
}
run 8gpu 8 "--steps 50 --warmup 10"
run 4gpu 4 "--steps 40 --warmup 5"
run 2gpu 2 "--steps 30 --warmup 5"
run kodak256_8gpu 8 "--workload kodak256 --steps 40 --warmup 5"
run tecnick_8gpu 8 "--workload tecnick --steps 30 --warmup 5"
run bb_8gpu 8 "--workload bb --steps 40 --warmup 5"
timeout 400 python -m pytest tests/test_gpu_cli.py -q -m gpu -k two_ranks > gpurun_out/pytest_r2_2gpu_cli.log 2>&1; tail -2 gpurun_out/pytest_r2_2gpu_cli.log
timeout 500 $TR --nproc-per-node 8 --master-port 29650 tools/tecnick_sweep.py --its 500 --out gpurun_out/tecnick_sweep_8gpu.json > gpurun_out/tecnick_sweep_8gpu.log 2>&1; tail -17 gpurun_out/tecnick_sweep_8gpu.log | cut -c1-200
