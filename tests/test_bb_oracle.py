"""CPU checks of the bits-back oracle (oracle/bb_oracle.py): the hand-derived pieces the device kernels implement
against autograd, and the invariants of the density model."""
import numpy as np
import torch

import sga_b200
from oracle import bb_oracle as B
from oracle import sga_oracle as O


def _prior(C, seed, dt):
    W = O._to(sga_b200.weights.make_synthetic_bb(C, seed), dt)
    return W, B.prior_params(W)


def test_prior_pdf_slope_equals_autograd():
    """Second-order forward mode through the CDF network (what k_zbb_path evaluates) == autograd of
    learned_prior.py:296-347's density."""
    W, eb = _prior(6, 3, torch.float64)
    z = (torch.randn(5, 2, 3, 6, dtype=torch.float64) * 4).requires_grad_(True)
    pdf = B.prior_pdf(z, eb)
    (g,) = torch.autograd.grad(pdf.sum(), [z])
    pdf_h, dpdf_h = B.prior_logpdf_slope(z.detach(), eb)
    assert torch.allclose(pdf_h, pdf.detach(), rtol=1e-10, atol=1e-14)
    assert torch.allclose(dpdf_h, g, rtol=1e-8, atol=1e-12)


def test_prior_pdf_integrates_to_one():
    W, eb = _prior(4, 1, torch.float64)
    grid = torch.linspace(-400, 400, 400001, dtype=torch.float64)
    z = grid[:, None].expand(-1, 4).contiguous()
    pdf = B.prior_pdf(z, eb)
    integral = torch.trapezoid(pdf, grid, dim=0)
    assert torch.all(pdf >= 0)
    assert torch.allclose(integral, torch.ones(4, dtype=torch.float64), atol=2e-3)


def test_bits_back_gradient_identities():
    """log q(z_tilde) with the reparameterised sample: d/d mean = 0, d/d logvar = -1/2 (bb_sga.py:104-107)."""
    mean = torch.randn(3, 4, dtype=torch.float64, requires_grad=True)
    lv = torch.randn(3, 4, dtype=torch.float64, requires_grad=True)
    eps = torch.randn(3, 4, dtype=torch.float64)
    zt = eps * torch.exp(0.5 * lv) + mean
    log_q = -0.5 * ((zt - mean) ** 2 * torch.exp(-lv) + lv + B.LOG2PI)
    gm, gl = torch.autograd.grad(log_q.sum(), [mean, lv])
    assert gm.abs().max() < 1e-12
    assert torch.allclose(gl, torch.full_like(gl, -0.5), atol=1e-12)


def test_bb_graph_runs_and_rate_only_matches_rd_with_zero_lambda():
    """The two sess.run signatures of bb_sga.py (:212-216 and :250-252) on a tiny model: the rate-only gradients equal
    the z-gradients of the full graph evaluated at the same y_tilde with lambda = 0."""
    C, n, h, w = 8, 1, 32, 32
    W = O._to(sga_b200.weights.make_synthetic_bb(C, 2), torch.float64)
    x = torch.from_numpy(sga_b200.synth.synthetic_images(n, h, w, 2).astype(np.float64) / 255.0)
    Wm = {k: v for k, v in W.items()}
    p = "analysis_transform/layer_%d/"
    t = x
    for i in range(4):
        t = O.corr_down(t, W[(p % i) + "kernel"], W[(p % i) + "bias"], 2)
        if i < 3:
            t = O.gdn(t, W[(p % i) + "gdn_%d/beta" % i], W[(p % i) + "gdn_%d/gamma" % i])
    y = t
    zm, zl = B.hyper_analysis(y, Wm)
    assert zm.shape == (n, 1, 1, C) and zl.shape == zm.shape
    eps = torch.randn(zm.shape, dtype=torch.float64)
    yt = torch.round(y)
    r = B.bb_graph(x, yt, zm, zl, W, 0.5, None, eps, 0.0, rate_only=True)
    g = torch.Generator().manual_seed(0)
    u = torch.rand((2,) + tuple(y.shape), generator=g, dtype=torch.float64).clamp(1e-6, 1 - 1e-6)
    full = B.bb_graph(x, yt, zm, zl, W, 0.5, (u[0], u[1]), eps, 0.0)      # integer y: the soft round is the identity
    assert torch.allclose(full["train_bpp"], r["train_bpp"], rtol=1e-10)
    assert torch.allclose(full["grad_z_mean"], r["grad_z_mean"], rtol=1e-8, atol=1e-12)
    assert torch.allclose(full["grad_z_logvar"], r["grad_z_logvar"], rtol=1e-8, atol=1e-12)
    ev = B.bb_eval(x, yt, zm, zl, eps, W)
    assert torch.allclose(ev["est_bpp"], r["y_bpp"] + r["z_bpp"] - r["bpp_back"], rtol=1e-10)


def test_golden_bb_fixture_reproduces():
    """tests/golden/bb_c16.npz (made by make_golden.py in the authoring container) pins the bits-back oracle."""
    import os
    import sys
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, gdir)
    import make_golden as G
    gold = dict(np.load(os.path.join(gdir, "bb_c16.npz")))
    here = G.HERE
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        G.HERE = tmp
        try:
            G.make_bb()
        finally:
            G.HERE = here
        new = dict(np.load(os.path.join(tmp, "bb_c16.npz")))
    for k in ("rd_loss", "train_bpp", "rate", "eval_est_bpp", "eval_psnr", "grad_y0", "grad_z_logvar0"):
        assert np.allclose(new[k], gold[k], rtol=2e-4, atol=1e-7), k
