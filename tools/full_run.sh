#!/bin/bash
# Full-size end-to-end run of the drop-in CLI on the GPU box: synthetic checkpoint + the 24x768x512 batch, 2000 SGA steps.
# usage: tools/full_run.sh [its]     (writes gpurun_out/full_run.log)
ITS=${1:-2000}
mkdir -p gpurun_out /tmp/sga_full/ckpt /tmp/sga_full/res
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import sga_b200
rn = sga_b200.weights.runname(192, 0.01)
sga_b200.weights.save_checkpoint("/tmp/sga_full/ckpt", rn, sga_b200.weights.make_synthetic(192, 0))
np.save("/tmp/sga_full/kodak_like.npy", sga_b200.synth.synthetic_images(24, 768, 512, 0))
print(rn)
PY
RN="mbt2018-num_filters=192-lmbda=0.01"
( time python sga.py --num_filters 192 --checkpoint_dir /tmp/sga_full/ckpt compress --sga_its $ITS --results_dir /tmp/sga_full/res $RN /tmp/sga_full/kodak_like.npy ) > gpurun_out/full_run.log 2>&1
tail -12 gpurun_out/full_run.log
