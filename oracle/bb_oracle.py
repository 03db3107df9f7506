"""CPU oracle for the bits-back variant (reference bb_sga.py) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Same rules as oracle/sga_oracle.py: only tests/ (and smoke / the bench CPU legs) may import this; parity is
**unpinned** at the tensorflow / tensorflow-compression boundary (PARITY_ASSUMPTIONS.md).

What bb_sga.py changes relative to sga.py (file:line into /root/reference):
  * z is not soft-rounded: z_tilde = eps * exp(z_logvar/2) + z_mean with eps ~ N(0, I)            bb_sga.py:99-105
  * its rate is the *density* of the flexible prior, p(z_tilde) = d CDF / dz of the same per-channel network
    (learned_prior.py:296-347), lower-bounded at 1e-9                                              bb_sga.py:109-111
  * the bits-back term  log q(z_tilde | mean, logvar)  (utils.log_normal_pdf, utils.py:75-77) is subtracted:
    eval_bpp = y_bpp + z_bpp - bpp_back                                                             bb_sga.py:132-141
  * gradients w.r.t. [y, z_mean, z_logvar] (bb_sga.py:156) and, for the second loop with y_tilde fixed to
    round(y), of train_bpp w.r.t. [z_mean, z_logvar] (bb_sga.py:157, 236-256)
  * h_a has 2C output channels, split into (z_mean_init, z_logvar_init), and is applied to y_tilde  bb_sga.py:69, 99
Everything downstream of (y_tilde, z_tilde) is the graph of sga.py and is reused from sga_oracle.
"""
import math

import numpy as np
import torch

from . import philox
from . import sga_oracle as O

TENSOR_ID_EPS = 3
LOG2PI = math.log(2.0 * math.pi)
PRIOR = "bmsh_j2018prior/"          # Keras snake-case of BMSHJ2018Prior


def prior_params(W):
    return {k.split("/")[-1]: v for k, v in W.items() if k.startswith(PRIOR)}


def normal_from_uniforms(u0, u1):
    """Box-Muller, float32: eps = sqrt(-2 ln u0) * cos(2 pi u1)  (the device kernel does the same)."""
    u0 = np.asarray(u0, dtype=np.float32)
    u1 = np.asarray(u1, dtype=np.float32)
    r = np.sqrt(np.float32(-2.0) * np.log(u0)).astype(np.float32)
    return (r * np.cos(np.float32(2.0 * math.pi) * u1)).astype(np.float32)


def draw_eps(z_mean, seed, step):
    u0, u1 = philox.sga_uniforms(z_mean.numel(), seed, step, TENSOR_ID_EPS)
    return torch.from_numpy(normal_from_uniforms(u0, u1)).reshape(z_mean.shape)


def hyper_analysis(y_tilde, W):
    """(z_mean_init, z_logvar_init) = split(h_a(y_tilde))   bb_sga.py:99, nn_models.py:73-103 with 2C outputs."""
    p = "hyper_analysis_transform/layer_%d/"
    t = torch.relu(O.corr_down(y_tilde, W[(p % 0) + "kernel"], W[(p % 0) + "bias"], 1))
    t = torch.relu(O.corr_down(t, W[(p % 1) + "kernel"], W[(p % 1) + "bias"], 2))
    t = O.corr_down(t, W[(p % 2) + "kernel"], None, 2)
    C = t.shape[-1] // 2
    return t[..., :C].contiguous(), t[..., C:].contiguous()


def prior_pdf(zt, eb):
    """p(z) = sigmoid'(L(z)) * L'(z)   (learned_prior.py:296-347: product of the layer Jacobians)."""
    L, dL = O.eb_logits_and_slope(zt, eb, want_slope=True)
    c = torch.sigmoid(L)
    return c * (1 - c) * dL


class _LowerBound(torch.autograd.Function):
    """math_ops.lower_bound (math_ops.py:63-76): identity-if-towards gradient."""

    @staticmethod
    def forward(ctx, inp, bound):
        ctx.save_for_backward(inp)
        ctx.bound = bound
        return torch.clamp(inp, min=bound)

    @staticmethod
    def backward(ctx, g):
        (inp,) = ctx.saved_tensors
        return g * ((inp >= ctx.bound) | (g < 0)).to(g.dtype), None


def bb_forward(x, yt, z_mean, z_logvar, eps, W, lmbda, n_ref=None, rate_only=False):
    """The bb_sga.py graph from (y_tilde, z_mean, z_logvar, eps) on, differentiable by autograd.
    rate_only: the second loop (bb_sga.py:236-256) -- no synthesis transform, loss = train_bpp."""
    N = yt.shape[0]
    n_ref = N if n_ref is None else n_ref
    H, Wd = (x.shape[1], x.shape[2]) if x is not None else (None, None)
    C = yt.shape[-1]
    zt = eps * torch.exp(0.5 * z_logvar) + z_mean                                              # bb_sga.py:104-105
    log_q = -0.5 * ((zt - z_mean) ** 2 * torch.exp(-z_logvar) + z_logvar + LOG2PI)             # utils.py:75-77
    pz = _LowerBound.apply(prior_pdf(zt, prior_params(W)), O.LIKELIHOOD_BOUND)                 # bb_sga.py:110-111
    h2, _ = O.hyper_synthesis_fwd(zt, W)
    mu = h2[:, :yt.shape[1], :yt.shape[2], :C]
    sigma = _LowerBound.apply(torch.exp(h2[:, :yt.shape[1], :yt.shape[2], C:]), O.SCALES_MIN)
    v = torch.abs(yt - mu)
    py = _LowerBound.apply(O.std_cumulative((0.5 - v) / sigma) - O.std_cumulative((-0.5 - v) / sigma),
                           O.LIKELIHOOD_BOUND)
    return zt, log_q, pz, py


def bb_graph(x, y, z_mean, z_logvar, W, T, noise_y, eps, lmbda, n_ref=None, rate_only=False):
    """One evaluation of the per-iteration sess.run of bb_sga.py:212-216 (or :250-252 when rate_only, where `y` is the
    fixed y_tilde = round(y)).  Returns losses, per-image terms and the gradients w.r.t. (y, z_mean, z_logvar)."""
    N, H, Wd, _ = x.shape
    n_ref = N if n_ref is None else n_ref
    npix = H * Wd
    z_mean = z_mean.clone().requires_grad_(True)
    z_logvar = z_logvar.clone().requires_grad_(True)
    if rate_only:
        yt = y
    else:
        y = y.clone().requires_grad_(True)
        yt = O.soft_round_autograd(y, noise_y[0], noise_y[1], T)
    zt, log_q, pz, py = bb_forward(x, yt, z_mean, z_logvar, eps, W, lmbda, n_ref)
    y_bpp = (-torch.log(py)).sum(dim=(1, 2, 3)) / (O.LN2 * npix)                                # bb_sga.py:136-137
    z_bpp = (-torch.log(pz)).sum(dim=(1, 2, 3)) / (O.LN2 * npix)                                # :138-139
    bpp_back = -log_q.sum(dim=(1, 2, 3)) / (O.LN2 * npix)                                       # :134-135
    eval_bpp = y_bpp + z_bpp - bpp_back                                                         # :140
    train_bpp = eval_bpp.sum() / n_ref                                                          # :141
    out = dict(train_bpp=train_bpp.detach(), y_bpp=y_bpp.detach(), z_bpp=z_bpp.detach(), bpp_back=bpp_back.detach(),
               z_tilde=zt.detach())
    if rate_only:
        g_mean, g_lv = torch.autograd.grad(train_bpp, [z_mean, z_logvar])                       # :157
        out.update(rd_loss=train_bpp.detach(), grad_z_mean=g_mean, grad_z_logvar=g_lv)
        return out
    xt, _ = O.synthesis_fwd(yt, W)
    xt = xt[:, :H, :Wd, :]
    train_mse = ((x - xt) ** 2).sum() / (n_ref * npix * 3) * 255.0 ** 2                         # :144-148
    rd_loss = lmbda * train_mse + train_bpp if lmbda > 0 else train_bpp                         # :153-156
    gy, g_mean, g_lv = torch.autograd.grad(rd_loss, [y, z_mean, z_logvar])                      # :156
    mse_r, psnr = O.psnr_rounded(x, xt.detach())
    out.update(rd_loss=rd_loss.detach(), train_mse=train_mse.detach(), psnr=psnr, mse_rounded=mse_r, grad_y=gy,
               grad_z_mean=g_mean, grad_z_logvar=g_lv, y_tilde=yt.detach())
    return out


def bb_eval(x, y_tilde, z_mean, z_logvar, eps, W):
    """The eval_tensors of bb_sga.py:176-177, 271-272 for a fixed y_tilde (= round(y))."""
    with torch.no_grad():
        N, H, Wd, _ = x.shape
        npix = H * Wd
        zt, log_q, pz, py = bb_forward(x, y_tilde, z_mean, z_logvar, eps, W, 0.0)
        y_bpp = (-torch.log(py)).sum(dim=(1, 2, 3)) / (O.LN2 * npix)
        z_bpp = (-torch.log(pz)).sum(dim=(1, 2, 3)) / (O.LN2 * npix)
        bpp_back = -log_q.sum(dim=(1, 2, 3)) / (O.LN2 * npix)
        xt, _ = O.synthesis_fwd(y_tilde, W)
        mse_r, psnr = O.psnr_rounded(x, xt[:, :H, :Wd, :])
        return dict(mse=mse_r, psnr=psnr, est_bpp=y_bpp + z_bpp - bpp_back, est_y_bpp=y_bpp, est_z_bpp=z_bpp,
                    est_bpp_back=bpp_back)


def prior_logpdf_slope(zt, eb):
    """Hand-derived d pdf / dz used by the device kernel: second-order forward mode through the CDF network.
    pdf = s(1-s) L',  d pdf/dz = s(1-s) [ (1-2s) L'^2 + L'' ],  s = sigmoid(L).  Returns (pdf, dpdf)."""
    dt = zt.dtype
    shape = zt.shape
    C = shape[-1]
    u = zt.reshape(-1, C)                                       # (B, C)
    h = [None] * 3
    d1 = [None] * 3
    d2 = [None] * 3
    sp = torch.nn.functional.softplus
    m0, b0, f0 = sp(eb["matrix_0"].to(dt)), eb["bias_0"].to(dt), torch.tanh(eb["factor_0"].to(dt))
    for r in range(3):
        a = m0[:, r, 0] * u + b0[:, r, 0]
        a1 = m0[:, r, 0].expand_as(u)
        a2 = torch.zeros_like(u)
        t = torch.tanh(a)
        gate = 1 + f0[:, r, 0] * (1 - t * t)
        h[r] = a + f0[:, r, 0] * t
        d2[r] = a2 * gate + a1 * a1 * f0[:, r, 0] * (-2 * t * (1 - t * t))
        d1[r] = a1 * gate
    for layer in (1, 2):
        m, b, f = sp(eb["matrix_%d" % layer].to(dt)), eb["bias_%d" % layer].to(dt), torch.tanh(eb["factor_%d" % layer].to(dt))
        nh, n1, n2 = [None] * 3, [None] * 3, [None] * 3
        for r in range(3):
            a = b[:, r, 0] + sum(m[:, r, k] * h[k] for k in range(3))
            a1 = sum(m[:, r, k] * d1[k] for k in range(3))
            a2 = sum(m[:, r, k] * d2[k] for k in range(3))
            t = torch.tanh(a)
            gate = 1 + f[:, r, 0] * (1 - t * t)
            nh[r] = a + f[:, r, 0] * t
            n2[r] = a2 * gate + a1 * a1 * f[:, r, 0] * (-2 * t * (1 - t * t))
            n1[r] = a1 * gate
        h, d1, d2 = nh, n1, n2
    m3, b3 = sp(eb["matrix_3"].to(dt)), eb["bias_3"].to(dt)
    L = b3[:, 0, 0] + sum(m3[:, 0, k] * h[k] for k in range(3))
    L1 = sum(m3[:, 0, k] * d1[k] for k in range(3))
    L2 = sum(m3[:, 0, k] * d2[k] for k in range(3))
    s = torch.sigmoid(L)
    pdf = s * (1 - s) * L1
    dpdf = s * (1 - s) * ((1 - 2 * s) * L1 * L1 + L2)
    return pdf.reshape(shape), dpdf.reshape(shape)
