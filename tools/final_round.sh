#!/bin/bash
# Final single-GPU evidence of a round: GPU test suite, smoke(), the default bench line, 2000-step bench lines of every
# workload (the metric is quoted on the 2000-step loop, sga.py:192), and the full drop-in CLI run.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu > gpurun_out/pytest_${TAG}_final.log 2>&1; tail -2 gpurun_out/pytest_${TAG}_final.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; tail -4 gpurun_out/smoke_${TAG}.log | cut -c1-200
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}_default.log 2>&1; grep '^{' gpurun_out/bench_${TAG}_default.log | tail -1 | cut -c1-160
for wl in kodak192 kodak256 tecnick bb; do
  timeout 500 python bench.py --workload $wl --steps 2000 --warmup 10 --no-cpu > gpurun_out/bench_${TAG}_${wl}_2000.log 2>&1
  grep '^{' gpurun_out/bench_${TAG}_${wl}_2000.log | tail -1 | cut -c1-160
done
tools/full_run.sh 2000
