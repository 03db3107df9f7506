#!/bin/bash
# A/B of the layer-3 kernel switches: SGA_L3F_REALLOC (forward: setmaxnreg register re-allocation), SGA_L3D_PAIR (fused dgrad:
# lane-pair stores).  Parity of the C = 192 gradient tests under every setting, then the bench's per-kernel times.
mkdir -p gpurun_out
i=0
for envs in "SGA_L3F_REALLOC=0 SGA_L3D_PAIR=0" "SGA_L3F_REALLOC=1 SGA_L3D_PAIR=1" "SGA_L3F_REALLOC=0 SGA_L3D_PAIR=0" "SGA_L3F_REALLOC=1 SGA_L3D_PAIR=1"; do
  i=$((i+1))
  if [ $i -le 2 ]; then
    env $envs timeout 300 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "test_grad_bf16_vs_oracle" > gpurun_out/l3ab_pytest_$i.log 2>&1
    echo "$envs: $(tail -1 gpurun_out/l3ab_pytest_$i.log | cut -c1-100)"
  fi
  env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/l3ab_bench_$i.log 2>&1
done
python tools/l3_times.py gpurun_out/l3ab_bench_*.log
