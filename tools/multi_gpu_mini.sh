#!/bin/bash
# 8-GPU lines of the two Kodak workloads with the current build (gpurun --gpus 8)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_r2c_8gpu.log 2>&1
grep '^{' gpurun_out/bench_r2c_8gpu.log | tail -1 | cut -c1-160
timeout 400 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --workload kodak256 --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_r2c_kodak256_8gpu.log 2>&1
grep '^{' gpurun_out/bench_r2c_kodak256_8gpu.log | tail -1 | cut -c1-160
