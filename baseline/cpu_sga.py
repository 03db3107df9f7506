"""bench.py's reference arm: the reference's SGA loop on the host CPU -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT.

The reference (TensorFlow 1.15 + tensorflow-compression 1.3, /root/reference/sga.py) cannot be installed in this image
(Python 3.12, no network), so this is the "PyTorch-CPU restatement of the reference TF1 graph" of BASELINE.md section 4,
driven exactly like the reference loop (sga.py:200-247):

  * the graph of sga.py:88-163 written forward-only (oracle.sga_graph_autograd) and differentiated w.r.t. [y, z] by
    autograd, like tf.gradients (sga.py:164);
  * every iteration re-feeds the host arrays x, y, z (sga.py:212-214) and fetches the gradients and the scalars;
  * host Adam over [y, z] with float32 state (adam.py:20-59; oracle.Adam is pinned to the reference's adam.py by
    tests/golden/adam_ref.npz -- adam.py itself lives under /root/reference, which does not exist on the GPU box);
  * Gumbel noise from a seeded torch generator (the reference draws it inside the graph with TF's Philox; any stream has
    the same cost class).

The whole reference batch is processed every iteration (24 images at the headline shape).  To bound host memory the
autograd graph is evaluated over `chunk` images at a time -- images are independent, every per-image gradient carries
1/n_ref (sga.py:147,150), so the arithmetic is the same as one 24-image sess.run.

Only bench.py (cpu_baseline / --impl reference) and tests/ import this module.
"""
import os
import time

import numpy as np
import torch

from oracle import sga_oracle as O


def run(x_u8, W, lmbda, steps, warmup, chunk=4, threads=None, seed=0, annealing_rate=1e-3, t0=700, lr=0.005,
        want_eval=False):
    """x_u8: uint8 [N,H,W,3].  Returns dict(seconds=[per timed iteration], cores, last={rd_loss, train_bpp, train_mse},
    eval=...)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    Wt = O._to(W, torch.float32)
    x_np = (x_u8.astype(np.float32) / 255.0)                       # sga.py:52-53
    n = x_np.shape[0]
    with torch.no_grad():
        parts = [O.analysis(torch.from_numpy(x_np[i:i + chunk]), Wt) for i in range(0, n, chunk)]   # sga.py:207
    y_cur = torch.cat([p[0] for p in parts]).numpy()
    z_cur = torch.cat([p[1] for p in parts]).numpy()
    opt = O.Adam(lr=lr)                                            # sga.py:208
    gen = torch.Generator().manual_seed(seed)
    secs, last = [], None
    for it in range(warmup + steps):
        t_start = time.perf_counter()
        T = O.annealed_temperature(it, r=annealing_rate, ub=0.5, scheme="exp0", t0=t0)     # sga.py:211
        gy, gz = np.empty_like(y_cur), np.empty_like(z_cur)
        rd = 0.0
        for i in range(0, n, chunk):                               # one sess.run's worth of work, `chunk` images at a time
            x = torch.from_numpy(x_np[i:i + chunk])                # x is fed every iteration (sga.py:212-214)
            y = torch.from_numpy(y_cur[i:i + chunk])
            z = torch.from_numpy(z_cur[i:i + chunk])
            noise = tuple(torch.rand(s.shape, generator=gen).clamp_(1e-7, 1.0) for s in (y, y, z, z))
            out = O.sga_graph_autograd(x, y, z, Wt, T, noise, lmbda, n_ref=n)
            gy[i:i + chunk] = out["grad_y"].numpy()
            gz[i:i + chunk] = out["grad_z"].numpy()
            rd += float(out["rd_loss"])
        new = opt.update([torch.from_numpy(y_cur), torch.from_numpy(z_cur)], [torch.from_numpy(gy), torch.from_numpy(gz)])
        y_cur, z_cur = new[0].numpy(), new[1].numpy()              # sga.py:215
        last = dict(rd_loss=rd)
        if it >= warmup:
            secs.append(time.perf_counter() - t_start)
    res = dict(seconds=secs, cores=threads, last=last, chunk=chunk)
    if want_eval:                                                  # sga.py:240-247
        with torch.no_grad():
            ev = [O.eval_hard(torch.from_numpy(x_np[i:i + chunk]), torch.from_numpy(y_cur[i:i + chunk]),
                              torch.from_numpy(z_cur[i:i + chunk]), Wt, lmbda) for i in range(0, n, chunk)]
        res["eval"] = {k: torch.cat([e[k] for e in ev]).numpy() for k in ev[0]}
    return res
