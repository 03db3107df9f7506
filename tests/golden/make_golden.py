"""Generates the committed golden fixtures.  Run HERE (the authoring container), where /root/reference exists:

    python tests/golden/make_golden.py

adam_ref.npz   -- trajectories of the reference's own numpy Adam (/root/reference/adam.py, imported unmodified).
                  The reference pins numpy 1.17.4 (requirements.txt:10), under which `float64 scalar * float32 array`
                  stays float32; numpy >= 2 (NEP 50) would silently promote to float64.  The import below swaps the
                  module-global `np` of adam.py for a shim whose scalar sqrt/power return Python floats ("weak"
                  scalars), which restores the numpy-1.17 float32 arithmetic without touching the reference code.
sga_c16.npz    -- the oracle's (oracle/sga_oracle.py) outputs on a tiny seeded case: per-step loss / gradients and the
                  final hard-rounded evaluation.  Pins the oracle against accidental edits and gives the CUDA fp32
                  path a fixed target that does not need the oracle at run time.
bb_c16.npz     -- the same for the bits-back variant (oracle/bb_oracle.py): 3 R-D steps over [y, z_mean, z_logvar], 3 rate
                  steps with y_tilde = round(y), the final evaluation (bb_sga.py:200-273), Philox + Box-Muller noise.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def reference_adam_class():
    sys.path.insert(0, "/root/reference")
    import adam as ref_adam  # noqa: the reference's file, unmodified
    shim = types.SimpleNamespace(
        sqrt=lambda x: float(np.sqrt(x)) if np.ndim(x) == 0 else np.sqrt(x),
        power=lambda a, b: float(np.power(a, b)) if np.ndim(a) == 0 and np.ndim(b) == 0 else np.power(a, b),
        zeros_like=np.zeros_like, square=np.square)
    ref_adam.np = shim
    return ref_adam.Adam


def make_adam():
    Adam = reference_adam_class()
    rng = np.random.default_rng(0)
    p0 = (rng.standard_normal(4099) * 3).astype(np.float32)
    steps = 12
    grads = (rng.standard_normal((steps, 4099)) * np.logspace(-6, 1, 4099)[None, :]).astype(np.float32)
    grads[:, :5] = 0.0
    opt = Adam(lr=0.005)
    p = [p0.copy()]
    out = []
    for it in range(steps):
        p = opt.update(p, [grads[it]])
        assert p[0].dtype == np.float32, p[0].dtype
        out.append(p[0].copy())
    np.savez_compressed(os.path.join(HERE, "adam_ref.npz"), p0=p0, grads=grads, params=np.stack(out))


def make_sga():
    import sga_b200
    from oracle import sga_oracle as O
    C, n, h, w, its, lmbda, seed = 16, 2, 48, 32, 4, 0.01, 5
    W = sga_b200.weights.make_synthetic(C, seed)
    x = torch.from_numpy(sga_b200.synth.synthetic_images(n, h, w, seed).astype(np.float32) / 255.0)
    rec = []
    # t0=1 so that the temperature actually anneals inside 4 steps
    y, z, ev = O.sga_run(x, W, lmbda, its, seed=seed, annealing_rate=0.2, t0=1, record=rec)
    np.savez_compressed(
        os.path.join(HERE, "sga_c16.npz"), C=C, n=n, h=h, w=w, its=its, lmbda=lmbda, seed=seed, annealing_rate=0.2, t0=1,
        rd_loss=np.array([float(r["rd_loss"]) for r in rec]), train_mse=np.array([float(r["train_mse"]) for r in rec]),
        train_bpp=np.array([float(r["train_bpp"]) for r in rec]), psnr=np.stack([r["psnr"].numpy() for r in rec]),
        grad_y0=rec[0]["grad_y"].numpy(), grad_z0=rec[0]["grad_z"].numpy(), T=np.array([r["T"] for r in rec]),
        y_final=y.numpy(), z_final=z.numpy(), **{"eval_" + k: v.numpy() for k, v in ev.items()})


def make_bb():
    import sga_b200
    from oracle import bb_oracle as B
    from oracle import sga_oracle as O
    C, n, h, w, its, r_its, lmbda, seed = 16, 2, 64, 48, 3, 3, 0.01, 7
    W = sga_b200.weights.make_synthetic_bb(C, seed)
    Wt = O._to(W, torch.float32)
    x = torch.from_numpy(sga_b200.synth.synthetic_images(n, h, w, seed).astype(np.float32) / 255.0)
    Wa = dict(Wt)
    Wa["hyper_analysis_transform/layer_2/kernel"] = Wt["hyper_analysis_transform/layer_2/kernel"][..., :C]
    y, _ = O.analysis(x, Wa)                                              # y_init (bb_sga.py:203)
    zm, zl = B.hyper_analysis(y, Wt)                                      # bb_sga.py:204
    opt = O.Adam(lr=0.005)
    rd, bpp = [], []
    g0 = None
    for it in range(its):
        T = O.annealed_temperature(it, r=0.2, ub=0.5, scheme="exp0", t0=1)
        uy = O.draw_noise(y, y[..., :1], seed, it)[:2]
        out = B.bb_graph(x, y, zm, zl, Wt, T, uy, B.draw_eps(zm, seed, it), lmbda)
        if it == 0:
            g0 = out
        rd.append(float(out["rd_loss"])); bpp.append(float(out["train_bpp"]))
        y, zm, zl = opt.update([y, zm, zl], [out["grad_y"], out["grad_z_mean"], out["grad_z_logvar"]])
    yt = torch.round(y)
    zm, zl = B.hyper_analysis(yt, Wt)                                     # bb_sga.py:247
    opt = O.Adam(lr=0.003)
    rate = []
    for it in range(r_its):
        out = B.bb_graph(x, yt, zm, zl, Wt, 1.0, None, B.draw_eps(zm, seed, its + it), lmbda, rate_only=True)
        rate.append(float(out["train_bpp"]))
        zm, zl = opt.update([zm, zl], [out["grad_z_mean"], out["grad_z_logvar"]])
    ev = B.bb_eval(x, yt, zm, zl, B.draw_eps(zm, seed, its + r_its), Wt)
    np.savez_compressed(
        os.path.join(HERE, "bb_c16.npz"), C=C, n=n, h=h, w=w, its=its, r_its=r_its, lmbda=lmbda, seed=seed, annealing_rate=0.2,
        t0=1, rd_loss=np.array(rd), train_bpp=np.array(bpp), rate=np.array(rate), grad_y0=g0["grad_y"].numpy(),
        grad_z_mean0=g0["grad_z_mean"].numpy(), grad_z_logvar0=g0["grad_z_logvar"].numpy(), y_tilde=yt.numpy(),
        z_mean_final=zm.numpy(), z_logvar_final=zl.numpy(), **{"eval_" + k: v.numpy() for k, v in ev.items()})


if __name__ == "__main__":
    make_adam()
    make_sga()
    make_bb()
    print("golden fixtures written to", HERE)
