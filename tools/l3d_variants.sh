#!/bin/bash
# A/B of the fused layer-3 dgrad variants (env SGA_L3D_VARIANT, bit 0 keep in TMEM, bit 1 pair stores, bit 2 register
# re-allocation): parity of the C = 192 gradient tests, then the bench's per-kernel time.
mkdir -p gpurun_out
for v in ${@:-0 1 2 3 4 5 6 7}; do
  SGA_L3D_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "test_grad_bf16_vs_oracle and 192" > gpurun_out/l3dv_pytest_$v.log 2>&1
  echo "variant $v: $(tail -1 gpurun_out/l3dv_pytest_$v.log | cut -c1-100)"
  SGA_L3D_VARIANT=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/l3dv_bench_$v.log 2>&1
done
python tools/l3_times.py gpurun_out/l3dv_bench_*.log
