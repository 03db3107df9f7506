"""CPU oracle for the SGA hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (the C-ABI library libsga_b200.so and its Python host) never
routes through it and fails loudly when the CUDA library is missing.

PARITY STATUS: **parity unpinned** at the third-party boundary.  The reference
(/root/reference/sga.py) delegates most arithmetic to tensorflow-compression==1.3,
tensorflow-probability==0.7.0 and tensorflow==1.15.0 (requirements.txt:21-24); none is vendored,
none is installable here, and the reference ships no tests or golden vectors.  What *is* pinned:
  * Adam against the reference's own adam.py (importable, numpy only)   -> tests/golden/adam_*.npz
  * the temperature schedule against utils.annealed_temperature semantics (utils.py:151-180)
  * every hand-derived backward against torch.autograd in float64
  * (de)conv index conventions against an independent direct-loop implementation
Every recalled third-party behaviour is listed in PARITY_ASSUMPTIONS.md.

All tensors are NHWC torch CPU tensors.  `dtype` float32 mirrors the reference placeholders
(sga.py:66,86,88,111); float64 is used for gradient checks.

Reference citations are file:line into /root/reference/.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import philox

LIKELIHOOD_BOUND = 1e-9      # tfc EntropyBottleneck / GaussianConditional default likelihood_bound
SCALES_MIN = 0.11            # sga.py:24 ; scale_bound defaults to scale_table[0]
EPSILON = 1e-5               # sga.py:30
TENSOR_ID_Y = 1
TENSOR_ID_Z = 2
LN2 = math.log(2.0)


# --------------------------------------------------------------------------------------
# a1. temperature schedule -- utils.py:151-180 (scheme 'exp0'), called at sga.py:211
# --------------------------------------------------------------------------------------
def annealed_temperature(t, r=1e-3, ub=0.5, lb=1e-8, scheme="exp0", t0=700):
    if scheme == "exp":
        tau = math.exp(-r * t)
    elif scheme == "exp0":
        tau = ub * math.exp(-r * (t - t0))
    elif scheme == "linear":
        tau = -r * (t - t0) + ub
    else:
        raise NotImplementedError(scheme)
    return min(max(tau, lb), ub)


# --------------------------------------------------------------------------------------
# a2. Gumbel-softmax soft rounding -- sga.py:88-98 (z) and :111-121 (y)
# --------------------------------------------------------------------------------------
def soft_round(v, u0, u1, T, eps=EPSILON):
    """Returns (v_tilde, dvtilde_dv).

    logits l0 = -atanh(clip(v-floor v))/T, l1 = -atanh(clip(ceil v - v))/T   (sga.py:91-93)
    tfp RelaxedOneHotCategorical(T, logits).sample(): G=-log(-log U); softmax((l+G)/T)
    (T divides twice, cf. the comment at sga.py:96);  v_tilde = floor*s0 + ceil*s1 (sga.py:98).
    floor/ceil carry no gradient; clip_by_value passes gradient only inside [-1+eps, 1-eps].
    """
    dt = v.dtype
    f = torch.floor(v)
    c = torch.ceil(v)
    lo = torch.tensor(-1 + eps, dtype=dt)
    hi = torch.tensor(1 - eps, dtype=dt)
    d = v - f
    e = c - v
    dc = torch.clamp(d, lo, hi)
    ec = torch.clamp(e, lo, hi)
    l0 = -torch.atanh(dc) / T
    l1 = -torch.atanh(ec) / T
    g0 = -torch.log(-torch.log(u0.to(dt)))
    g1 = -torch.log(-torch.log(u1.to(dt)))
    a = ((l1 + g1) - (l0 + g0)) / T
    s1 = torch.sigmoid(a)
    s0 = 1 - s1
    vt = f * s0 + c * s1
    # d vt / d v = (c - f) * s1 (1 - s1) / T * (dl1/dv - dl0/dv)
    pass0 = ((d >= lo) & (d <= hi)).to(dt)
    pass1 = ((e >= lo) & (e <= hi)).to(dt)
    dl0 = -pass0 / ((1 - dc * dc) * T)       # d l0 / d v
    dl1 = pass1 / ((1 - ec * ec) * T)        # d l1 / d v   (d e / d v = -1)
    dvt = (c - f) * s1 * s0 * (dl1 - dl0) / T
    return vt, dvt


def soft_round_autograd(v, u0, u1, T, eps=EPSILON):
    """Same forward written op-for-op like sga.py:88-98, differentiable by autograd (tests)."""
    f = torch.floor(v).detach()
    c = torch.ceil(v).detach()
    l0 = -torch.atanh(torch.clamp(v - f, -1 + eps, 1 - eps)) / T
    l1 = -torch.atanh(torch.clamp(c - v, -1 + eps, 1 - eps)) / T
    logits = torch.stack([l0, l1], dim=-1)
    g = -torch.log(-torch.log(torch.stack([u0, u1], dim=-1).to(v.dtype)))
    sample = torch.softmax((logits + g) / T, dim=-1)
    return (torch.stack([f, c], dim=-1) * sample).sum(-1)


# --------------------------------------------------------------------------------------
# a3. factorised prior likelihood -- tfc.EntropyBottleneck._likelihood (sga.py:100-104);
#     CDF network restated in-repo at learned_prior.py:78-121, forward-mode derivative at :296-347
# --------------------------------------------------------------------------------------
def eb_logits_and_slope(u, eb, want_slope=True):
    """u: (..., C).  eb: dict with matrix_i (C,d_{i+1},d_i), bias_i (C,d_{i+1},1), factor_i (C,d_{i+1},1).
    Returns L(u) and dL/du, both (..., C)."""
    dt = u.dtype
    shape = u.shape
    C = shape[-1]
    x = u.reshape(-1, C).t().reshape(C, 1, -1)             # (C,1,B)  learned_prior.py:148-149
    jac = None
    nlayer = 4
    for i in range(nlayer):
        m = F.softplus(eb["matrix_%d" % i].to(dt))          # learned_prior.py:99-104
        x = torch.matmul(m, x) + eb["bias_%d" % i].to(dt)   # :105-110
        if i < nlayer - 1:
            fac = torch.tanh(eb["factor_%d" % i].to(dt))    # :112-119
            th = torch.tanh(x)
            x = x + fac * th
            nl_grad = 1 + fac * (1 - th * th)               # :321-324
        else:
            nl_grad = torch.ones_like(x)
        if want_slope:
            # J_i = diag(nl_grad) @ M_i ; chain: J = J_i @ J  (learned_prior.py:334-341)
            ji = nl_grad.permute(2, 0, 1).unsqueeze(-1) * m                # (B,C,r,d)
            jac = ji if jac is None else torch.matmul(ji, jac)
    L = x.reshape(C, -1).t().reshape(shape)
    if want_slope:
        dL = jac[..., 0, 0].reshape(shape)
        return L, dL
    return L, None


def eb_likelihood(zt, eb):
    """Returns (p, dp/dzt) with p = |sigmoid(s*upper) - sigmoid(s*lower)|, s = -sign(lower+upper)
    (stop-gradient), lower/upper = L(zt -/+ 0.5)."""
    lower, dlower = eb_logits_and_slope(zt - 0.5, eb)
    upper, dupper = eb_logits_and_slope(zt + 0.5, eb)
    s = -torch.sign(lower + upper)
    su = torch.sigmoid(s * upper)
    sl = torch.sigmoid(s * lower)
    diff = su - sl
    p = torch.abs(diff)
    dp = torch.sign(diff) * s * (su * (1 - su) * dupper - sl * (1 - sl) * dlower)
    return p, dp


# --------------------------------------------------------------------------------------
# a6. Gaussian conditional likelihood -- tfc.GaussianConditional._likelihood (sga.py:129-136),
#     restated in-repo at utils.py:80-102
# --------------------------------------------------------------------------------------
def std_cumulative(x):
    return 0.5 * torch.special.erfc(-(2 ** -0.5) * x)     # utils.py:83


def gauss_likelihood(yt, mu, logsigma):
    """Returns p and the partials (dp/dyt, dp/dmu, dp/dsigma_bounded, sigma, pass_mask_ge)."""
    dt = yt.dtype
    sigma = torch.exp(logsigma)                            # sga.py:108
    sb = torch.clamp(sigma, min=SCALES_MIN)                # lower_bound(scale, scale_table[0])
    d = yt - mu
    v = torch.abs(d)                                       # utils.py:96
    a = (0.5 - v) / sb
    b = (-0.5 - v) / sb
    p = std_cumulative(a) - std_cumulative(b)              # utils.py:98-101
    inv_sqrt_2pi = 1.0 / math.sqrt(2 * math.pi)
    pa = torch.exp(-0.5 * a * a) * inv_sqrt_2pi
    pb = torch.exp(-0.5 * b * b) * inv_sqrt_2pi
    sg = torch.sign(d)
    dp_dy = -sg * (pa - pb) / sb
    dp_dsb = -(a * pa - b * pb) / sb
    return p, dp_dy, -dp_dy, dp_dsb, sigma, (sigma >= SCALES_MIN)


# --------------------------------------------------------------------------------------
# a4/a7/a14. transforms -- nn_models.py; tfc.SignalConv2D / tfc.GDN semantics (see header)
# --------------------------------------------------------------------------------------
def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def deconv_s2(x, k, b=None):
    """SignalConv2D(corr=False, strides_up=2, same_zeros), k=(kh,kw,Cin,Cout):
    out[t] = sum_i x[i] k[t+2-2i], output exactly 2x input (nn_models.py:48-63,152-159)."""
    w = k.permute(2, 3, 0, 1)
    return _nhwc(F.conv_transpose2d(_nchw(x), w, b, stride=2, padding=k.shape[0] // 2, output_padding=1))


def deconv_s2_dgrad(g, k):
    """Adjoint of deconv_s2 w.r.t. x: stride-2 correlation of g with the un-flipped kernel."""
    w = k.permute(2, 3, 0, 1)          # read as (out=Cin, in=Cout, kh, kw)
    return _nhwc(F.conv2d(_nchw(g), w, None, stride=2, padding=k.shape[0] // 2))


def conv_s1_true(x, k, b=None):
    """SignalConv2D(corr=False, stride 1, same_zeros): true convolution (nn_models.py:160-163)."""
    w = k.permute(2, 3, 0, 1)
    return _nhwc(F.conv_transpose2d(_nchw(x), w, b, stride=1, padding=k.shape[0] // 2))


def conv_s1_true_dgrad(g, k):
    w = k.permute(2, 3, 0, 1)
    return _nhwc(F.conv2d(_nchw(g), w, None, stride=1, padding=k.shape[0] // 2))


def corr_down(x, k, b=None, stride=2):
    """SignalConv2D(corr=True, strides_down=stride, same_zeros) (nn_models.py:14-28,85-96)."""
    w = k.permute(3, 2, 0, 1)
    return _nhwc(F.conv2d(_nchw(x), w, b, stride=stride, padding=k.shape[0] // 2))


def gdn_norm(x, beta, gamma):
    """norm_i = beta_i + sum_j gamma[j,i] x_j^2 (1x1 conv, gamma as [in=j,out=i])."""
    return torch.matmul(x * x, gamma) + beta


def igdn(x, beta, gamma):
    s = torch.sqrt(gdn_norm(x, beta, gamma))
    return x * s, s


def igdn_bwd(g, x, s, gamma):
    """g_x = g*sqrt(n) + x * ((g*x/sqrt(n)) @ gamma^T)."""
    t = g * x / s
    return g * s + x * torch.matmul(t, gamma.t())


def gdn(x, beta, gamma):
    return x * torch.rsqrt(gdn_norm(x, beta, gamma))


def analysis(x, W):
    """g_a (nn_models.py:5-36) and h_a (:73-103): y_init, z_init (sga.py:77-78)."""
    p = "analysis_transform/layer_%d/"
    t = x
    for i in range(4):
        t = corr_down(t, W[(p % i) + "kernel"], W[(p % i) + "bias"], 2)
        if i < 3:
            t = gdn(t, W[(p % i) + "gdn_%d/beta" % i], W[(p % i) + "gdn_%d/gamma" % i])
    y = t
    p = "hyper_analysis_transform/layer_%d/"
    t = torch.relu(corr_down(y, W[(p % 0) + "kernel"], W[(p % 0) + "bias"], 1))
    t = torch.relu(corr_down(t, W[(p % 1) + "kernel"], W[(p % 1) + "bias"], 2))
    z = corr_down(t, W[(p % 2) + "kernel"], None, 2)
    return y, z


HS = "mb_t2018hyper_synthesis_transform/layer_%d/"
GS = "synthesis_transform/layer_%d/"


def hyper_synthesis_fwd(zt, W):
    a0 = deconv_s2(zt, W[(HS % 0) + "kernel"], W[(HS % 0) + "bias"])
    h0 = torch.relu(a0)
    a1 = deconv_s2(h0, W[(HS % 1) + "kernel"], W[(HS % 1) + "bias"])
    h1 = torch.relu(a1)
    h2 = conv_s1_true(h1, W[(HS % 2) + "kernel"], W[(HS % 2) + "bias"])
    return h2, (a0, a1)


def hyper_synthesis_bwd(g2, saved, W):
    a0, a1 = saved
    g1 = conv_s1_true_dgrad(g2, W[(HS % 2) + "kernel"]) * (a1 > 0).to(g2.dtype)
    g0 = deconv_s2_dgrad(g1, W[(HS % 1) + "kernel"]) * (a0 > 0).to(g2.dtype)
    return deconv_s2_dgrad(g0, W[(HS % 0) + "kernel"])


def synthesis_fwd(yt, W):
    saved = []
    t = yt
    for i in range(3):
        xpre = deconv_s2(t, W[(GS % i) + "kernel"], W[(GS % i) + "bias"])
        t, s = igdn(xpre, W[(GS % i) + "igdn_%d/beta" % i], W[(GS % i) + "igdn_%d/gamma" % i])
        saved.append((xpre, s))
    xt = deconv_s2(t, W[(GS % 3) + "kernel"], W[(GS % 3) + "bias"])
    return xt, saved


def synthesis_bwd(gx, saved, W):
    g = deconv_s2_dgrad(gx, W[(GS % 3) + "kernel"])
    for i in (2, 1, 0):
        xpre, s = saved[i]
        g = igdn_bwd(g, xpre, s, W[(GS % i) + "igdn_%d/gamma" % i])
        g = deconv_s2_dgrad(g, W[(GS % i) + "kernel"])
    return g


# --------------------------------------------------------------------------------------
# the operator seam: (y, z, x, T) -> (grad_y, grad_z, rd_loss, train_mse, train_bpp, psnr[N])
# sga.py:212-214 ; graph sga.py:88-174
# --------------------------------------------------------------------------------------
def _to(W, dt):
    return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))).to(dt) for k, v in W.items()}


def eb_params(W):
    return {k.split("/")[-1]: v for k, v in W.items() if k.startswith("entropy_bottleneck/")}


def psnr_rounded(x, xt):
    """sga.py:167-174: x*255 vs round(clip(x_tilde,0,1)*255) (half-to-even), tf.image.psnr."""
    xr = torch.round(torch.clamp(xt, 0, 1) * 255)
    mse = ((x * 255 - xr) ** 2).mean(dim=(1, 2, 3))
    return mse, 20 * math.log10(255.0) - 10 * torch.log10(mse)


def forward_from_tilde(x, yt, zt, W, lmbda, n_ref=None):
    """Everything downstream of (y_tilde, z_tilde): shared by the SGA step and the final hard eval
    (sga.py:244-247 feeds y_tilde/z_tilde directly)."""
    N, H, Wd, _ = x.shape
    n_ref = N if n_ref is None else n_ref
    npix = H * Wd
    pz, dpz = eb_likelihood(zt, eb_params(W))
    pz_c = torch.clamp(pz, min=LIKELIHOOD_BOUND)
    h2, hs_saved = hyper_synthesis_fwd(zt, W)
    C = yt.shape[-1]
    Hy, Wy = yt.shape[1], yt.shape[2]
    mu = h2[:, :Hy, :Wy, :C]
    logsigma = h2[:, :Hy, :Wy, C:]
    py, dpy_dy, dpy_dmu, dpy_dsb, sigma, sig_ge = gauss_likelihood(yt, mu, logsigma)
    py_c = torch.clamp(py, min=LIKELIHOOD_BOUND)
    xt_full, gs_saved = synthesis_fwd(yt, W)
    xt = xt_full[:, :H, :Wd, :]
    y_bpp = (-torch.log(py_c)).sum(dim=(1, 2, 3)) / (LN2 * npix)      # sga.py:144
    z_bpp = (-torch.log(pz_c)).sum(dim=(1, 2, 3)) / (LN2 * npix)      # sga.py:145
    sse = ((x - xt) ** 2).sum(dim=(1, 2, 3))
    train_mse = sse.sum() / (n_ref * npix * 3) * 255.0 ** 2            # sga.py:150-154
    train_bpp = (y_bpp + z_bpp).sum() / n_ref                          # sga.py:147
    rd_loss = lmbda * train_mse + train_bpp if lmbda > 0 else train_bpp  # sga.py:160-163
    mse_r, psnr = psnr_rounded(x, xt)
    out = dict(rd_loss=rd_loss, train_mse=train_mse, train_bpp=train_bpp, psnr=psnr, mse_rounded=mse_r,
               y_bpp=y_bpp, z_bpp=z_bpp, sse=sse, x_tilde=xt, mu=mu, logsigma=logsigma)
    ctx = dict(pz_c=pz_c, dpz=dpz, py_c=py_c, dpy_dy=dpy_dy, dpy_dmu=dpy_dmu, dpy_dsb=dpy_dsb, sigma=sigma,
               sig_ge=sig_ge, hs_saved=hs_saved, gs_saved=gs_saved, xt_full=xt_full, h2_shape=h2.shape)
    return out, ctx


def backward_to_tilde(x, yt, zt, W, lmbda, ctx, n_ref=None):
    """d rd_loss / d (y_tilde, z_tilde)  (the part of tf.gradients, sga.py:164, below the soft round)."""
    N, H, Wd, _ = x.shape
    n_ref = N if n_ref is None else n_ref
    npix = H * Wd
    dt = x.dtype
    C = yt.shape[-1]
    Hy, Wy = yt.shape[1], yt.shape[2]
    rate_scale = 1.0 / (LN2 * npix * n_ref)
    # rate of y: d(-log p_c)/dp = -1/p_c ; the lower_bound gradient always passes because the
    # upstream gradient is negative (math_ops.py:74-76)
    gpy = -rate_scale / ctx["py_c"]
    g_yt = gpy * ctx["dpy_dy"]
    g_mu = gpy * ctx["dpy_dmu"]
    g_sb = gpy * ctx["dpy_dsb"]
    # lower_bound(sigma, 0.11): pass iff sigma >= bound or grad < 0 (math_ops.py:74-76)
    g_sigma = g_sb * (ctx["sig_ge"] | (g_sb < 0)).to(dt)
    g_logsigma = g_sigma * ctx["sigma"]
    # distortion
    if lmbda > 0:
        gx_full = torch.zeros_like(ctx["xt_full"])
        gx_full[:, :H, :Wd, :] = (2.0 * 255.0 ** 2 * lmbda / (n_ref * npix * 3)) * (ctx["xt_full"][:, :H, :Wd, :] - x)
        g_yt = g_yt + synthesis_bwd(gx_full, ctx["gs_saved"], W)
    # hyper path
    g_h2 = torch.zeros(ctx["h2_shape"], dtype=dt)
    g_h2[:, :Hy, :Wy, :C] = g_mu
    g_h2[:, :Hy, :Wy, C:] = g_logsigma
    g_zt = hyper_synthesis_bwd(g_h2, ctx["hs_saved"], W)
    g_zt = g_zt + (-rate_scale / ctx["pz_c"]) * ctx["dpz"]
    return g_yt, g_zt


def sga_graph(x, y, z, W, T, noise, lmbda, n_ref=None):
    """One evaluation of the reference's per-iteration sess.run (sga.py:212-214).
    noise = (uy0, uy1, uz0, uz1) uniforms in (0,1) shaped like y / z."""
    uy0, uy1, uz0, uz1 = noise
    zt, dzt = soft_round(z, uz0, uz1, T)
    yt, dyt = soft_round(y, uy0, uy1, T)
    out, ctx = forward_from_tilde(x, yt, zt, W, lmbda, n_ref)
    g_yt, g_zt = backward_to_tilde(x, yt, zt, W, lmbda, ctx, n_ref)
    out["grad_y"] = g_yt * dyt
    out["grad_z"] = g_zt * dzt
    out["y_tilde"] = yt
    out["z_tilde"] = zt
    out["g_ytilde"] = g_yt
    out["g_ztilde"] = g_zt
    return out


def sga_graph_autograd(x, y, z, W, T, noise, lmbda, n_ref=None):
    """The same graph written forward-only and differentiated by torch.autograd: an independent
    check of every hand-derived backward above.  lower_bound's custom gradient is emulated with a
    straight-through clamp where the reference always passes it."""
    class LowerBound(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp, bound):
            ctx.save_for_backward(inp)
            ctx.bound = bound
            return torch.clamp(inp, min=bound)

        @staticmethod
        def backward(ctx, g):
            (inp,) = ctx.saved_tensors
            return g * ((inp >= ctx.bound) | (g < 0)).to(g.dtype), None

    N, H, Wd, _ = x.shape
    n_ref = N if n_ref is None else n_ref
    npix = H * Wd
    y = y.clone().requires_grad_(True)
    z = z.clone().requires_grad_(True)
    uy0, uy1, uz0, uz1 = noise
    zt = soft_round_autograd(z, uz0, uz1, T)
    yt = soft_round_autograd(y, uy0, uy1, T)
    eb = eb_params(W)
    lower, _ = eb_logits_and_slope(zt - 0.5, eb, want_slope=False)
    upper, _ = eb_logits_and_slope(zt + 0.5, eb, want_slope=False)
    s = -torch.sign(lower + upper).detach()
    pz = LowerBound.apply(torch.abs(torch.sigmoid(s * upper) - torch.sigmoid(s * lower)), LIKELIHOOD_BOUND)
    h2, _ = hyper_synthesis_fwd(zt, W)
    C = y.shape[-1]
    mu = h2[:, :y.shape[1], :y.shape[2], :C]
    sigma = LowerBound.apply(torch.exp(h2[:, :y.shape[1], :y.shape[2], C:]), SCALES_MIN)
    v = torch.abs(yt - mu)
    py = LowerBound.apply(std_cumulative((0.5 - v) / sigma) - std_cumulative((-0.5 - v) / sigma), LIKELIHOOD_BOUND)
    xt, _ = synthesis_fwd(yt, W)
    xt = xt[:, :H, :Wd, :]
    bpp = ((-torch.log(py)).sum() + (-torch.log(pz)).sum()) / (LN2 * npix * n_ref)
    mse = ((x - xt) ** 2).sum() / (n_ref * npix * 3) * 255.0 ** 2
    loss = lmbda * mse + bpp if lmbda > 0 else bpp
    gy, gz = torch.autograd.grad(loss, [y, z])
    return dict(rd_loss=loss.detach(), grad_y=gy, grad_z=gz)


# --------------------------------------------------------------------------------------
# a12. Adam -- adam.py:20-59 with float32 state (numpy 1.17 semantics)
# --------------------------------------------------------------------------------------
class Adam:
    def __init__(self, lr=0.005, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
        self.iterations = 0
        self.ms = None
        self.vs = None

    def lr_t(self):
        t = self.iterations + 1
        return self.lr * (math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t))   # adam.py:40-42

    def update(self, params, grads):
        if self.ms is None:
            self.ms = [torch.zeros_like(p) for p in params]
            self.vs = [torch.zeros_like(p) for p in params]
        lr_t = self.lr_t()
        out = []
        for i, (p, g) in enumerate(zip(params, grads)):
            dt = p.dtype
            m = (self.beta_1 * self.ms[i] + (1.0 - self.beta_1) * g).to(dt)        # adam.py:50
            v = (self.beta_2 * self.vs[i] + (1.0 - self.beta_2) * g * g).to(dt)    # adam.py:51
            out.append((p - lr_t * m / (torch.sqrt(v) + self.epsilon)).to(dt))     # adam.py:52
            self.ms[i], self.vs[i] = m, v
        self.iterations += 1
        return out


# --------------------------------------------------------------------------------------
# the loop -- sga.py:200-247
# --------------------------------------------------------------------------------------
def draw_noise(y, z, seed, step):
    uy0, uy1 = philox.sga_uniforms(y.numel(), seed, step, TENSOR_ID_Y)
    uz0, uz1 = philox.sga_uniforms(z.numel(), seed, step, TENSOR_ID_Z)
    f = lambda a, ref: torch.from_numpy(a).reshape(ref.shape)
    return f(uy0, y), f(uy1, y), f(uz0, z), f(uz1, z)


def eval_hard(x, y, z, W, lmbda):
    """sga.py:240-247: hard-rounded latents (np.round = half-to-even) fed as y_tilde/z_tilde."""
    out, _ = forward_from_tilde(x, torch.round(y), torch.round(z), W, lmbda)
    return dict(mse=out["mse_rounded"], psnr=out["psnr"], est_bpp=out["y_bpp"] + out["z_bpp"],
                est_y_bpp=out["y_bpp"], est_z_bpp=out["z_bpp"])


def sga_run(x, W, lmbda, its, seed=0, lr=0.005, annealing_rate=1e-3, t0=700, T_ub=0.5, dtype=torch.float32,
            n_ref=None, record=None):
    """x: (N,H,W,3) float in [0,1].  Runs init (sga.py:207), `its` SGA steps (sga.py:210-215) and the
    final hard evaluation (sga.py:240-247).  Returns (y, z, eval dict)."""
    W = _to(W, dtype)
    x = x.to(dtype)
    y, z = analysis(x, W)
    opt = Adam(lr=lr)
    for it in range(its):
        T = annealed_temperature(it, r=annealing_rate, ub=T_ub, scheme="exp0", t0=t0)
        out = sga_graph(x, y, z, W, T, draw_noise(y, z, seed, it), lmbda, n_ref)
        if record is not None:
            record.append({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()
                           if k in ("rd_loss", "train_mse", "train_bpp", "psnr", "grad_y", "grad_z")} | {"T": T})
        y, z = opt.update([y, z], [out["grad_y"], out["grad_z"]])
    return y, z, eval_hard(x, y, z, W, lmbda)
