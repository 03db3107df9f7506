"""Development aid: run N SGA steps at the headline shape (optionally with SGA_DBG wait accounting) and print the
per-kernel event times.  usage: [SGA_DBG=8] [SGA_NO_PAIR=1] python tools/dbg_step.py [steps] [n_images]"""
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
C, H, W = 192, 768, 512
Wt = sga_b200.weights.make_synthetic(C, 0)
x = torch.from_numpy(sga_b200.synth.synthetic_images(n, H, W, 0).astype(np.float32) / 255.0).cuda()
eng = SgaEngine(Wt, n, H, W, C, 0.01, n_ref=24, operand="bf16", seed=0)
y, z = eng.encode_init(x)
st = eng.new_state(y.clone(), z.clone())
torch.cuda.synchronize()
eng.profile(True)
for it in range(steps):
    eng.step(x, st, 0.5, it, want_stats=False)
torch.cuda.synchronize()
prof = eng.profile_read()
for k, v in sorted(prof.items()):
    print("%-12s %8.4f ms/step" % (k, v[0] / steps))
print("total %.4f ms/step" % (sum(v[0] for v in prof.values()) / steps))
