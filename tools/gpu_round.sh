#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, bench, ncu launch list and full captures of the hot kernels.
# usage: tools/gpu_round.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-x}; TESTS=${2:-tests}; NCU=${3:-ncu}
mkdir -p gpurun_out
if [ "$TESTS" = tests ]; then
  timeout 600 python tests/run_gpu_isolated.py --timeout 150 tests/test_gpu_conv.py tests/test_gpu_engine.py > gpurun_out/iso_$TAG.log 2>&1
  grep -v "^PASS" gpurun_out/iso_$TAG.log | tail -40
fi
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.log 2>&1; tail -3 gpurun_out/bench_$TAG.log
if [ "$NCU" = ncu ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l_$TAG.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_l3|k_igdn' -s 9 -c 3 -f -o gpurun_out/prof_l3_$TAG \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_a_$TAG.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' -s 36 -c 12 -f -o gpurun_out/prof_conv_$TAG \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_a_$TAG.log gpurun_out/ncu_b_$TAG.log
fi
