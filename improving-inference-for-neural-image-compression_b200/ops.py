"""Operator-level wrappers over the C ABI (torch CUDA tensors in, torch CUDA tensors out).  These call the same
kernels the SGA step uses and exist for the parity tests."""
import ctypes

import torch

from . import _lib
from .engine import _ptr, _stream

KIND = {"up2": 0, "up2_dgrad": 1, "s1_true": 2, "s1_true_dgrad": 3, "down2_corr": 4, "s1_corr": 5}
_OPERAND = {"f32": _lib.OPERAND_F32, "bf16": _lib.OPERAND_BF16}


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise _lib.SgaError("operators take contiguous float32 CUDA tensors")


def soft_round(v, temperature, u=None, clip_eps=1e-5, seed=0, step=0, tensor_id=1):
    _need_cuda(v, u)
    vt, dv = torch.empty_like(v), torch.empty_like(v)
    _lib.check(_lib.load().sga_op_soft_round(_ptr(v), _ptr(u), v.numel(), temperature, clip_eps, seed, step, tensor_id,
                                            _ptr(vt), _ptr(dv), _stream()))
    return vt, dv


def eb_likelihood(zt, matrices, biases, factors):
    _need_cuda(zt, *matrices, *biases, *factors)
    p, dp = torch.empty_like(zt), torch.empty_like(zt)
    vp = ctypes.c_void_p
    m = (vp * 4)(*[t.data_ptr() for t in matrices])
    b = (vp * 4)(*[t.data_ptr() for t in biases])
    f = (vp * 3)(*[t.data_ptr() for t in factors])
    _lib.check(_lib.load().sga_op_eb_likelihood(_ptr(zt), zt.numel(), zt.shape[-1], m, b, f, _ptr(p), _ptr(dp),
                                               _stream()))
    return p, dp


def gauss_likelihood(yt, mu, log_sigma):
    _need_cuda(yt, mu, log_sigma)
    p, dy, ds = torch.empty_like(yt), torch.empty_like(yt), torch.empty_like(yt)
    _lib.check(_lib.load().sga_op_gauss_likelihood(_ptr(yt), _ptr(mu), _ptr(log_sigma), yt.numel(), _ptr(p), _ptr(dy),
                                                  _ptr(ds), _stream()))
    return p, dy, ds


def adam_(p, m, v, g, lr, beta1, beta2, eps, iterations):
    _need_cuda(p, m, v, g)
    _lib.check(_lib.load().sga_op_adam(_ptr(p), _ptr(m), _ptr(v), _ptr(g), p.numel(), lr, beta1, beta2, eps, iterations,
                                      _stream()))


def conv(kind, x, kernel, bias=None, relu=False, operand="f32"):
    """kernel (kh,kw,Cin,Cout) in the reference layout; x NHWC."""
    _need_cuda(x, kernel, bias)
    n, h, w, cx = x.shape
    kh, kw, cin, cout = kernel.shape
    k = KIND[kind]
    dgrad = k in (1, 3)
    assert cx == (cout if dgrad else cin), "channel mismatch"
    if k == 0:
        ho, wo = 2 * h, 2 * w
    elif k == 1:
        ho, wo = h // 2, w // 2
    elif k == 4:
        ho, wo = (h + 1) // 2, (w + 1) // 2
    else:
        ho, wo = h, w
    out = torch.empty((n, ho, wo, cin if dgrad else cout), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().sga_op_conv(k, _OPERAND[operand], _ptr(x), _ptr(kernel), _ptr(bias), n, h, w, kh, kw, cin,
                                      cout, int(relu), _ptr(out), _stream()))
    return out


def gdn(x, beta, gamma, inverse=True, operand="f32"):
    _need_cuda(x, beta, gamma)
    out, s = torch.empty_like(x), torch.empty_like(x)
    c = x.shape[-1]
    _lib.check(_lib.load().sga_op_gdn(_OPERAND[operand], int(inverse), _ptr(x), _ptr(beta), _ptr(gamma), x.numel() // c,
                                     c, _ptr(out), _ptr(s), _stream()))
    return out, s


def igdn_bwd(g, x, s, gamma, operand="f32"):
    _need_cuda(g, x, s, gamma)
    gx = torch.empty_like(x)
    c = x.shape[-1]
    _lib.check(_lib.load().sga_op_igdn_bwd(_OPERAND[operand], _ptr(g), _ptr(x), _ptr(s), _ptr(gamma), x.numel() // c, c,
                                          _ptr(gx), _stream()))
    return gx
