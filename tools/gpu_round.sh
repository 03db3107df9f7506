#!/bin/bash
# Runs on the GPU box (via gpurun): the evidence of a round -- launch list, full ncu capture of one step's tensor kernels.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc|k_conv_igdn|k_l3' -s 42 -c 14 -f \
    -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_a_$TAG.log 2>&1
tail -2 gpurun_out/ncu_l_$TAG.log gpurun_out/ncu_a_$TAG.log | cut -c1-200
