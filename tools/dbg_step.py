"""Development aid: run N SGA steps at the headline shape (optionally with SGA_DBG wait accounting) and print the
per-kernel event times.  usage: [SGA_DBG=8] [SGA_NO_PAIR=1] python tools/dbg_step.py [steps] [n_images] [C] [H] [W] [n_ref]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import sga_b200  # noqa: E402
from sga_b200.engine import SgaEngine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
C = int(sys.argv[3]) if len(sys.argv) > 3 else 192
H = int(sys.argv[4]) if len(sys.argv) > 4 else 768
W = int(sys.argv[5]) if len(sys.argv) > 5 else 512
n_ref = int(sys.argv[6]) if len(sys.argv) > 6 else 24
Wt = sga_b200.weights.make_synthetic(C, 0)
x = torch.from_numpy(sga_b200.synth.synthetic_images(n, H, W, 0).astype(np.float32) / 255.0).cuda()
eng = SgaEngine(Wt, n, H, W, C, 0.01 if C != 256 else 0.08, n_ref=n_ref, operand="bf16", seed=0)
y, z = eng.encode_init(x)
st = eng.new_state(y.clone(), z.clone())
for it in range(3):
    eng.step(x, st, 0.5, it, want_stats=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(steps):                      # un-instrumented: the number that matters
    eng.step(x, st, 0.5, it, want_stats=False)
e1.record()
torch.cuda.synchronize()
eng.profile(True)                            # per-kernel events (serialise the launches: breakdown only)
for it in range(steps):
    eng.step(x, st, 0.5, it, want_stats=False)
torch.cuda.synchronize()
print("shape n=%d %dx%d C=%d: %.4f ms/step wall (CUDA events), %.1f it/s of this shard" % (n, H, W, C, e0.elapsed_time(e1) / steps, 1000.0 * steps / e0.elapsed_time(e1)))
prof = eng.profile_read()
for k, v in sorted(prof.items()):
    print("%-12s %8.4f ms/step" % (k, v[0] / steps))
print("total %.4f ms/step" % (sum(v[0] for v in prof.values()) / steps))
