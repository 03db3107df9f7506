"""Parity of the bits-back variant (C ABI: sga_bb_hyper_init / sga_bb_grad / sga_bb_step / sga_bb_eval) against
oracle/bb_oracle.py (restatement of reference bb_sga.py) with identical noise."""
import numpy as np
import pytest
import torch

import sga_b200
from oracle import bb_oracle as B
from oracle import sga_oracle as O
from helpers import rel_err, rms_rel_err

pytestmark = pytest.mark.gpu


def _case(C, n, h, w, seed=0):
    W = sga_b200.weights.make_synthetic_bb(C, seed)
    Wt = O._to(W, torch.float32)
    x = torch.from_numpy(sga_b200.synth.synthetic_images(n, h, w, seed).astype(np.float32) / 255.0)
    p = "analysis_transform/layer_%d/"
    t = x
    for i in range(4):
        t = O.corr_down(t, Wt[(p % i) + "kernel"], Wt[(p % i) + "bias"], 2)
        if i < 3:
            t = O.gdn(t, Wt[(p % i) + "gdn_%d/beta" % i], Wt[(p % i) + "gdn_%d/gamma" % i])
    g = torch.Generator().manual_seed(seed + 1)
    y = (t + 0.3 * torch.randn(t.shape, generator=g)).contiguous()
    return W, Wt, x, y


def _engine(W, n, h, w, C, lmbda, operand, seed=0):
    from sga_b200.engine import BbEngine
    return BbEngine(W, n, h, w, C, lmbda, operand=operand, seed=seed)


@pytest.mark.parametrize("C,n,h,w", [(16, 2, 64, 64), (32, 1, 80, 48)])
def test_bb_init_f32(C, n, h, w):
    """y_init (bb_sga.py:203) and split(h_a(y_tilde)) (bb_sga.py:99, 204)."""
    W, Wt, x, y = _case(C, n, h, w)
    eng = _engine(W, n, h, w, C, 0.01, "f32")
    y_dev = eng.encode_init(x.cuda())
    y_o, _ = O.analysis(x, {**Wt, "hyper_analysis_transform/layer_2/kernel": Wt["hyper_analysis_transform/layer_2/kernel"][..., :C]})
    assert rel_err(y_dev, y_o) < 2e-5
    zm, zl = eng.hyper_init(y.cuda())
    zm_o, zl_o = B.hyper_analysis(y, Wt)
    assert tuple(zm.shape) == tuple(zm_o.shape)
    assert rel_err(zm, zm_o) < 2e-5 and rel_err(zl, zl_o) < 2e-5


@pytest.mark.parametrize("C,n,h,w,T,lmbda", [(16, 2, 64, 64, 0.5, 0.01), (32, 1, 80, 48, 0.2, 0.08)])
def test_bb_grad_f32(C, n, h, w, T, lmbda):
    """bb_sga.py:212-216 in float32 with injected noise: gradients w.r.t. y, z_mean, z_logvar and the statistics."""
    W, Wt, x, y = _case(C, n, h, w, seed=C)
    zm, zl = B.hyper_analysis(y, Wt)
    g = torch.Generator().manual_seed(7)
    u = torch.rand((2,) + tuple(y.shape), generator=g).clamp(1e-6, 1 - 1e-6)
    eps = torch.randn(zm.shape, generator=g)
    ref = B.bb_graph(x, y, zm, zl, Wt, T, (u[0], u[1]), eps, lmbda)
    eng = _engine(W, n, h, w, C, lmbda, "f32")
    gy, gm, gl, stats = eng.grad(x.cuda(), y.cuda(), zm.cuda(), zl.cuda(), T, 0, u.contiguous().cuda(), eps.cuda())
    s = eng.unpack_stats(stats.cpu().numpy())
    assert abs(s["rd_loss"] - float(ref["rd_loss"])) <= 3e-5 * abs(float(ref["rd_loss"]))
    assert abs(s["train_bpp"] - float(ref["train_bpp"])) <= 3e-5 * abs(float(ref["train_bpp"]))
    assert np.abs(s["est_bpp_back"] - ref["bpp_back"].numpy()).max() <= 1e-5 * np.abs(ref["bpp_back"].numpy()).max()
    assert np.abs(s["est_z_bpp"] - ref["z_bpp"].numpy()).max() <= 1e-4 * np.abs(ref["z_bpp"].numpy()).max()
    assert np.abs(s["psnr"] - ref["psnr"].numpy()).max() < 1e-3
    assert rel_err(gy, ref["grad_y"]) < 2e-4
    assert rel_err(gm, ref["grad_z_mean"]) < 3e-4 and rel_err(gl, ref["grad_z_logvar"]) < 3e-4


def test_bb_rate_only_and_philox_f32():
    """Second loop (bb_sga.py:250-252): y_tilde fixed, gradients of train_bpp; eps from Philox + Box-Muller."""
    C, n, h, w, seed, step = 16, 2, 64, 48, 5, 9
    W, Wt, x, y = _case(C, n, h, w, seed=3)
    yt = torch.round(y)
    zm, zl = B.hyper_analysis(yt, Wt)
    eps = B.draw_eps(zm, seed, step)
    ref = B.bb_graph(x, yt, zm, zl, Wt, 1.0, None, eps, 0.01, rate_only=True)
    eng = _engine(W, n, h, w, C, 0.01, "f32", seed=seed)
    _, gm, gl, stats = eng.grad(None, yt.cuda(), zm.cuda(), zl.cuda(), 1.0, step, rate_only=True)
    s = eng.unpack_stats(stats.cpu().numpy())
    assert abs(s["train_bpp"] - float(ref["train_bpp"])) <= 5e-5 * abs(float(ref["train_bpp"]))
    assert rel_err(gm, ref["grad_z_mean"]) < 5e-4 and rel_err(gl, ref["grad_z_logvar"]) < 5e-4
    ev = eng.eval(x.cuda(), yt.cuda(), zm.cuda(), zl.cuda(), step=step)
    ev_o = B.bb_eval(x, yt, zm, zl, eps, Wt)
    assert np.abs(ev["est_bpp"] - ev_o["est_bpp"].numpy()).max() < 1e-3      # north_star tolerance: 1e-3 bpp
    assert np.abs(ev["psnr"] - ev_o["psnr"].numpy()).max() < 0.01             # 0.01 dB


@pytest.mark.parametrize("C,n,h,w", [(64, 2, 128, 96), (192, 1, 128, 128)])
def test_bb_grad_bf16_vs_oracle(C, n, h, w):
    """bf16 tcgen05 path of the bits-back graph (both loops) vs the float32 oracle."""
    W, Wt, x, y = _case(C, n, h, w, seed=C + 2)
    zm, zl = B.hyper_analysis(y, Wt)
    g = torch.Generator().manual_seed(11)
    u = torch.rand((2,) + tuple(y.shape), generator=g).clamp(1e-6, 1 - 1e-6)
    eps = torch.randn(zm.shape, generator=g)
    ref = B.bb_graph(x, y, zm, zl, Wt, 0.35, (u[0], u[1]), eps, 0.01)
    eng = _engine(W, n, h, w, C, 0.01, "bf16")
    gy, gm, gl, stats = eng.grad(x.cuda(), y.cuda(), zm.cuda(), zl.cuda(), 0.35, 0, u.contiguous().cuda(), eps.cuda())
    s = eng.unpack_stats(stats.cpu().numpy())
    assert torch.isfinite(gy).all() and torch.isfinite(gm).all() and torch.isfinite(gl).all()
    assert abs(s["train_bpp"] - float(ref["train_bpp"])) <= 3e-3 * abs(float(ref["train_bpp"]))
    assert rms_rel_err(gy, ref["grad_y"]) < 3e-2
    assert rms_rel_err(gm, ref["grad_z_mean"]) < 0.1 and rms_rel_err(gl, ref["grad_z_logvar"]) < 0.1
    yt = torch.round(y)
    ref2 = B.bb_graph(x, yt, zm, zl, Wt, 1.0, None, eps, 0.01, rate_only=True)
    _, gm2, gl2, _ = eng.grad(None, yt.cuda(), zm.cuda(), zl.cuda(), 1.0, 0, eps_z=eps.cuda(), rate_only=True)
    assert rms_rel_err(gm2, ref2["grad_z_mean"]) < 0.1 and rms_rel_err(gl2, ref2["grad_z_logvar"]) < 0.1


def test_bb_loop_runs_and_reduces_rate():
    """The two loops of bb_sga.py:200-273 end to end (bf16): finite results, integer y_tilde; with the noise frozen the
    rate loop (Adam, lr 0.003) lowers train_bpp."""
    C, n, h, w = 64, 1, 96, 96
    W, Wt, x, _ = _case(C, n, h, w, seed=4)
    eng = _engine(W, n, h, w, C, 0.01, "bf16", seed=1)
    xd = x.cuda()
    y_tilde, zm, zl, ev = eng.run(xd, its=20, r_its=10, annealing_rate=0.05, t0=5)
    assert torch.equal(y_tilde, torch.round(y_tilde))
    assert all(np.isfinite(ev[k]).all() for k in ("psnr", "est_bpp", "est_y_bpp", "est_z_bpp", "est_bpp_back"))
    assert np.allclose(ev["est_bpp"], ev["est_y_bpp"] + ev["est_z_bpp"] - ev["est_bpp_back"], rtol=1e-5)
    zm0, zl0 = eng.hyper_init(y_tilde)
    eps = torch.randn(zm0.shape, generator=torch.Generator().manual_seed(3)).cuda()
    bpp0 = float(eng.grad(None, y_tilde, zm0, zl0, 1.0, 0, eps_z=eps, rate_only=True)[3][2])
    st = eng.new_state(y_tilde, zm0.clone(), zl0.clone())
    for it in range(60):
        eng.step(None, st, 1.0, it, 0.003, eps_z=eps, rate_only=True, want_stats=False)
    bpp1 = float(eng.grad(None, y_tilde, st["z_mean"], st["z_logvar"], 1.0, 0, eps_z=eps, rate_only=True)[3][2])
    assert bpp1 < bpp0, (bpp0, bpp1)


def test_golden_bb_fixture_f32():
    """Both loops of bb_sga.py:200-273 on the device in fp32 with the Philox / Box-Muller stream against the committed
    fixture tests/golden/bb_c16.npz (no oracle at run time)."""
    import os
    from sga_b200.schedule import annealed_temperature
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "bb_c16.npz"))
    C, n, h, w, its, r_its, seed = (int(gold[k]) for k in ("C", "n", "h", "w", "its", "r_its", "seed"))
    lmbda = float(gold["lmbda"])
    W = sga_b200.weights.make_synthetic_bb(C, seed)
    x = torch.from_numpy(sga_b200.synth.synthetic_images(n, h, w, seed).astype(np.float32) / 255.0).cuda()
    eng = _engine(W, n, h, w, C, lmbda, "f32", seed=seed)
    y = eng.encode_init(x)
    zm, zl = eng.hyper_init(y)
    gy, gm, gl, _ = eng.grad(x, y, zm, zl, annealed_temperature(0, r=0.2, ub=0.5, scheme="exp0", t0=1), 0)
    assert rel_err(gy, gold["grad_y0"]) < 3e-4 and rel_err(gm, gold["grad_z_mean0"]) < 5e-4
    assert rel_err(gl, gold["grad_z_logvar0"]) < 5e-4
    st = eng.new_state(y, zm, zl)
    for it in range(its):
        T = annealed_temperature(it, r=float(gold["annealing_rate"]), ub=0.5, scheme="exp0", t0=int(gold["t0"]))
        eng.step(x, st, T, it, 0.005)
        s = eng.unpack_stats(eng.stats.cpu().numpy())
        assert abs(s["rd_loss"] - gold["rd_loss"][it]) <= 2e-4 * abs(gold["rd_loss"][it]), it
    y_tilde = torch.round(st["y"])
    assert (y_tilde.cpu().numpy() != gold["y_tilde"]).mean() < 1e-3          # a rounding tie may flip an element
    zm, zl = eng.hyper_init(y_tilde)
    st2 = eng.new_state(y_tilde, zm, zl)
    for it in range(r_its):
        eng.step(None, st2, 1.0, its + it, 0.003, rate_only=True, adam_iterations=it)
        s = eng.unpack_stats(eng.stats.cpu().numpy())
        assert abs(s["train_bpp"] - gold["rate"][it]) <= 5e-4 * abs(gold["rate"][it]), it
    ev = eng.eval(x, y_tilde, st2["z_mean"], st2["z_logvar"], step=its + r_its)
    assert np.abs(ev["est_bpp"] - gold["eval_est_bpp"]).max() < 1e-3          # north_star tolerance: 1e-3 bpp
    assert np.abs(ev["psnr"] - gold["eval_psnr"]).max() < 0.01                # 0.01 dB
