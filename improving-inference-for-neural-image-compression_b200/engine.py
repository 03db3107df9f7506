"""Host driver of the SGA loop: owns the torch device tensors, calls the C-ABI library.

Mirrors the body of compress() in the reference (sga.py:200-253): init via the analysis transforms (sga.py:207),
`its` iterations of {temperature (sga.py:211) -> graph + gradients (sga.py:212-214) -> Adam (sga.py:215)}, then
the hard-rounded evaluation (sga.py:240-247).  Nothing crosses the host inside the loop except the temperature
scalar; statistics are fetched only at log points."""
import ctypes

import numpy as np
import torch

from . import _lib
from .schedule import annealed_temperature
from .weights import EB, GA, GS, HA, HS

EVAL_FIELDS = ["mse", "psnr", "msssim", "msssim_db", "est_bpp", "est_y_bpp", "est_z_bpp"]   # sga.py:183


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class SgaEngine:
    """One context = one GPU = one shard of images of a reference batch."""

    def __init__(self, weights, n_images, height, width, num_filters, lmbda, n_ref=None, operand="bf16", seed=0,
                 lr=0.005, beta1=0.9, beta2=0.999, adam_eps=1e-8, clip_eps=1e-5, device=None, first_image=0):
        if not torch.cuda.is_available():
            raise _lib.SgaError("sga_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        torch.cuda.set_device(self.device)
        self.cfg = _lib.SgaConfig(
            n_images=n_images, height=height, width=width, num_filters=num_filters,
            n_ref=n_images if n_ref is None else n_ref,
            operand={"f32": _lib.OPERAND_F32, "bf16": _lib.OPERAND_BF16}[operand],
            lmbda=lmbda, lr=lr, beta1=beta1, beta2=beta2, adam_eps=adam_eps, clip_eps=clip_eps, seed=seed, first_image=first_image)
        self.n, self.h, self.w, self.c = n_images, height, width, num_filters
        self.ctx = ctypes.c_void_p()
        _lib.check(self.lib.sga_create(ctypes.byref(self.cfg), ctypes.byref(self.ctx)))
        shp = [ctypes.c_int32() for _ in range(4)]
        _lib.check(self.lib.sga_latent_shape(self.ctx, *[ctypes.byref(s) for s in shp]))
        self.hy, self.wy, self.hz, self.wz = [s.value for s in shp]
        self.y_shape = (n_images, self.hy, self.wy, num_filters)
        self.z_shape = (n_images, self.hz, self.wz, num_filters)
        self._bind(weights)
        self.n_stats = _lib.STATS_HEADER + _lib.STATS_PER_IMAGE * n_images
        self.stats = torch.zeros(self.n_stats, dtype=torch.float32, device=self.device)
        self.launches = 0

    # ---- weights -----------------------------------------------------------------------------------
    def _bind(self, W):
        dev = self.device
        self._w = {k: torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float32).to(dev).contiguous()
                   for k, v in W.items()}
        w = _lib.SgaWeights()

        def P(name):
            return self._w[name].data_ptr()
        for i in range(4):
            w.ga_kernel[i] = P((GA % i) + "kernel"); w.ga_bias[i] = P((GA % i) + "bias")
            w.gs_kernel[i] = P((GS % i) + "kernel"); w.gs_bias[i] = P((GS % i) + "bias")
            w.eb_matrix[i] = P(EB + "matrix_%d" % i); w.eb_bias[i] = P(EB + "bias_%d" % i)
        for i in range(3):
            w.ga_beta[i] = P((GA % i) + "gdn_%d/beta" % i); w.ga_gamma[i] = P((GA % i) + "gdn_%d/gamma" % i)
            w.gs_beta[i] = P((GS % i) + "igdn_%d/beta" % i); w.gs_gamma[i] = P((GS % i) + "igdn_%d/gamma" % i)
            w.ha_kernel[i] = P((HA % i) + "kernel")
            w.hs_kernel[i] = P((HS % i) + "kernel"); w.hs_bias[i] = P((HS % i) + "bias")
            w.eb_factor[i] = P(EB + "factor_%d" % i)
        for i in range(2):
            w.ha_bias[i] = P((HA % i) + "bias")
        _lib.check(self.lib.sga_bind_weights(self.ctx, ctypes.byref(w), _stream()))

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.sga_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self):
        return int(self.lib.sga_workspace_bytes(self.ctx))

    # ---- the operator seam -------------------------------------------------------------------------
    def encode_init(self, x):
        """(y_init, z_init) = (g_a(x), h_a(g_a(x)))  -- sga.py:77-78, 207."""
        y = torch.empty(self.y_shape, dtype=torch.float32, device=self.device)
        z = torch.empty(self.z_shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.sga_encode_init(self.ctx, _ptr(x), _ptr(y), _ptr(z), _stream()))
        return y, z

    def grad(self, x, y, z, temperature, step, noise_y=None, noise_z=None):
        """The reference's sess.run([rd_gradients, ...]) (sga.py:212-214). Returns (grad_y, grad_z, stats tensor)."""
        gy = torch.empty_like(y)
        gz = torch.empty_like(z)
        _lib.check(self.lib.sga_grad(self.ctx, _ptr(x), _ptr(y), _ptr(z), float(temperature), int(step), _ptr(noise_y),
                                     _ptr(noise_z), _ptr(gy), _ptr(gz), _ptr(self.stats), _stream()))
        return gy, gz, self.stats

    def new_state(self, y, z):
        return dict(y=y, z=z, m_y=torch.zeros_like(y), v_y=torch.zeros_like(y), m_z=torch.zeros_like(z),
                    v_z=torch.zeros_like(z))

    def step(self, x, st, temperature, step, noise_y=None, noise_z=None, want_stats=True):
        """One SGA iteration in place on the state dict (sga.py:211-215)."""
        _lib.check(self.lib.sga_step(self.ctx, _ptr(x), _ptr(st["y"]), _ptr(st["z"]), _ptr(st["m_y"]), _ptr(st["v_y"]),
                                     _ptr(st["m_z"]), _ptr(st["v_z"]), float(temperature), int(step), _ptr(noise_y),
                                     _ptr(noise_z), _ptr(self.stats) if want_stats else None, _stream()))

    def eval_hard(self, x, y, z):
        """sga.py:240-247 -- hard-rounded latents. Returns dict of per-image numpy arrays."""
        _lib.check(self.lib.sga_eval_hard(self.ctx, _ptr(x), _ptr(y), _ptr(z), _ptr(self.stats), _stream()))
        return self.unpack_stats(self.stats.cpu().numpy())

    def reconstruction(self):
        """x_tilde [n,H,W,3] (float, un-clipped) of the most recent fp32 forward (e.g. eval_hard)."""
        xt = torch.empty((self.n, self.h, self.w, 3), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.sga_reconstruction(self.ctx, _ptr(xt), _stream()))
        return xt

    def profile(self, enable):
        _lib.check(self.lib.sga_profile_enable(self.ctx, int(bool(enable))))

    def profile_read(self):
        """{label name: (milliseconds, launch groups)} accumulated since the last read."""
        ms = (ctypes.c_float * _lib.PROF_LABELS)()
        cnt = (ctypes.c_int32 * _lib.PROF_LABELS)()
        _lib.check(self.lib.sga_profile_read(self.ctx, ms, cnt, _lib.PROF_LABELS))
        return {_lib.PROF_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(_lib.PROF_LABELS)
                if i in _lib.PROF_NAMES and cnt[i] > 0}

    def launch_count(self):
        return int(self.lib.sga_launch_count())

    def unpack_stats(self, s):
        per = s[_lib.STATS_HEADER:].reshape(self.n, _lib.STATS_PER_IMAGE)
        return dict(rd_loss=float(s[0]), train_mse=float(s[1]), train_bpp=float(s[2]), psnr=per[:, 0].copy(),
                    est_y_bpp=per[:, 1].copy(), est_z_bpp=per[:, 2].copy(), est_bpp=per[:, 1] + per[:, 2],
                    sse=per[:, 3].copy(), mse=per[:, 4].copy())

    # ---- the loop ----------------------------------------------------------------------------------
    def run(self, x, its=2000, annealing_rate=1e-3, t0=700, T_ub=0.5, log_itv=100, log=None, verbose=False):
        """x: float32 [n,H,W,3] in [0,1] on the device.  Returns (y, z, eval dict)."""
        y, z = self.encode_init(x)
        st = self.new_state(y, z)
        for it in range(its):
            T = annealed_temperature(it, r=annealing_rate, ub=T_ub, scheme="exp0", t0=t0)       # sga.py:211
            at_log = log is not None and (it % log_itv == 0 or it + 1 == its)                    # sga.py:216
            self.step(x, st, T, it, want_stats=at_log)
            if at_log:
                s = self.unpack_stats(self.stats.cpu().numpy())
                extra = self.eval_hard(x, st["y"], st["z"]) if verbose else None
                log(it, T, s, extra)
        return st["y"], st["z"], self.eval_hard(x, st["y"], st["z"])
