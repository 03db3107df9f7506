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
from .weights import EB, GA, GS, HA, HS, PRIOR

EVAL_FIELDS = ["mse", "psnr", "msssim", "msssim_db", "est_bpp", "est_y_bpp", "est_z_bpp"]   # sga.py:183


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class SgaEngine:
    """One context = one GPU = one shard of images of a reference batch."""

    VARIANT = _lib.VARIANT_SGA
    PRIOR_PREFIX = EB
    STATS_PER_IMAGE = _lib.STATS_PER_IMAGE

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
            lmbda=lmbda, lr=lr, beta1=beta1, beta2=beta2, adam_eps=adam_eps, clip_eps=clip_eps, seed=seed,
            first_image=first_image, variant=self.VARIANT)
        self.n, self.h, self.w, self.c = n_images, height, width, num_filters
        self.ctx = ctypes.c_void_p()
        _lib.check(self.lib.sga_create(ctypes.byref(self.cfg), ctypes.byref(self.ctx)))
        shp = [ctypes.c_int32() for _ in range(4)]
        _lib.check(self.lib.sga_latent_shape(self.ctx, *[ctypes.byref(s) for s in shp]))
        self.hy, self.wy, self.hz, self.wz = [s.value for s in shp]
        self.y_shape = (n_images, self.hy, self.wy, num_filters)
        self.z_shape = (n_images, self.hz, self.wz, num_filters)
        self._bind(weights)
        self.n_stats = _lib.STATS_HEADER + self.STATS_PER_IMAGE * n_images
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
            w.eb_matrix[i] = P(self.PRIOR_PREFIX + "matrix_%d" % i); w.eb_bias[i] = P(self.PRIOR_PREFIX + "bias_%d" % i)
        for i in range(3):
            w.ga_beta[i] = P((GA % i) + "gdn_%d/beta" % i); w.ga_gamma[i] = P((GA % i) + "gdn_%d/gamma" % i)
            w.gs_beta[i] = P((GS % i) + "igdn_%d/beta" % i); w.gs_gamma[i] = P((GS % i) + "igdn_%d/gamma" % i)
            w.ha_kernel[i] = P((HA % i) + "kernel")
            w.hs_kernel[i] = P((HS % i) + "kernel"); w.hs_bias[i] = P((HS % i) + "bias")
            w.eb_factor[i] = P(self.PRIOR_PREFIX + "factor_%d" % i)
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

    def loop(self, x, st, temperatures, first_step, want_stats=True):
        """len(temperatures) consecutive SGA iterations in place on the state dict, device-resident (graph replay on bf16
        contexts): iteration k uses temperatures[k] and step first_step + k.  self.stats then holds the LAST iteration's
        statistics."""
        n = len(temperatures)
        arr = (ctypes.c_float * max(n, 1))(*[float(t) for t in temperatures])
        _lib.check(self.lib.sga_loop(self.ctx, _ptr(x), _ptr(st["y"]), _ptr(st["z"]), _ptr(st["m_y"]), _ptr(st["v_y"]),
                                     _ptr(st["m_z"]), _ptr(st["v_z"]), arr, int(first_step), n,
                                     _ptr(self.stats) if want_stats else None, _stream()))

    def eval_hard(self, x, y, z):
        """sga.py:240-247 -- hard-rounded latents. Returns dict of per-image numpy arrays."""
        _lib.check(self.lib.sga_eval_hard(self.ctx, _ptr(x), _ptr(y), _ptr(z), _ptr(self.stats), _stream()))
        return self.unpack_stats(self.stats.cpu().numpy())

    def eval_tilde(self, x, y_tilde, z_tilde):
        """Everything downstream of given (y_tilde, z_tilde), fp32, nothing rounded (mbt2018.py:171-185 on the baseline's
        mean-centred y_hat).  x = None: reconstruction only (the decoder's synthesis, mbt2018.py:275-283)."""
        _lib.check(self.lib.sga_eval_tilde(self.ctx, _ptr(x), _ptr(y_tilde), _ptr(z_tilde), _ptr(self.stats), _stream()))
        return None if x is None else self.unpack_stats(self.stats.cpu().numpy())

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
        temps = [annealed_temperature(it, r=annealing_rate, ub=T_ub, scheme="exp0", t0=t0) for it in range(its)]  # sga.py:211
        # iterations run in device-resident segments that end at the reference's log points (sga.py:216)
        ends = [it for it in range(its) if it % log_itv == 0 or it + 1 == its] if log is not None else ([its - 1] if its else [])
        first = 0
        for end in ends:
            self.loop(x, st, temps[first:end + 1], first, want_stats=log is not None)
            if log is not None:
                s = self.unpack_stats(self.stats.cpu().numpy())
                extra = self.eval_hard(x, st["y"], st["z"]) if verbose else None
                log(end, temps[end], s, extra)
            first = end + 1
        return st["y"], st["z"], self.eval_hard(x, st["y"], st["z"])


BB_EVAL_FIELDS = EVAL_FIELDS + ["est_bpp_back"]                                                   # bb_sga.py:176


class BbEngine(SgaEngine):
    """The bits-back variant (reference bb_sga.py): same transforms, z_tilde = eps*exp(z_logvar/2) + z_mean with the
    density of the flexible prior as its rate and the bits-back term log q subtracted (bb_sga.py:99-141)."""

    VARIANT = _lib.VARIANT_BITS_BACK
    PRIOR_PREFIX = PRIOR
    STATS_PER_IMAGE = _lib.BB_STATS_PER_IMAGE

    def encode_init(self, x):
        """y_init = g_a(x)  (bb_sga.py:83, 203)."""
        y = torch.empty(self.y_shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.sga_encode_init(self.ctx, _ptr(x), _ptr(y), None, _stream()))
        return y

    def hyper_init(self, y_tilde):
        """(z_mean_init, z_logvar_init) = split(h_a(y_tilde))  (bb_sga.py:99, 204, 247)."""
        zm = torch.empty(self.z_shape, dtype=torch.float32, device=self.device)
        zl = torch.empty(self.z_shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.sga_bb_hyper_init(self.ctx, _ptr(y_tilde), _ptr(zm), _ptr(zl), _stream()))
        return zm, zl

    def grad(self, x, y, z_mean, z_logvar, temperature, step, noise_y=None, eps_z=None, rate_only=False):
        """bb_sga.py:212-216 (or :250-252 with rate_only, where y is the fixed y_tilde).
        Returns (grad_y or None, grad_z_mean, grad_z_logvar, stats tensor)."""
        gy = None if rate_only else torch.empty_like(y)
        gm, gl = torch.empty_like(z_mean), torch.empty_like(z_logvar)
        _lib.check(self.lib.sga_bb_grad(self.ctx, _ptr(x), _ptr(y), _ptr(z_mean), _ptr(z_logvar), float(temperature),
                                        int(step), _ptr(noise_y), _ptr(eps_z), int(bool(rate_only)), _ptr(gy), _ptr(gm),
                                        _ptr(gl), _ptr(self.stats), _stream()))
        return gy, gm, gl, self.stats

    def new_state(self, y, z_mean, z_logvar):
        st = dict(y=y, z_mean=z_mean, z_logvar=z_logvar)
        for k in ("y", "z_mean", "z_logvar"):
            st["m_" + k] = torch.zeros_like(st[k])
            st["v_" + k] = torch.zeros_like(st[k])
        return st

    def step(self, x, st, temperature, step, lr, noise_y=None, eps_z=None, rate_only=False, want_stats=True,
             adam_iterations=None):
        """One iteration in place on the state dict; `step` keys the noise, `adam_iterations` (default: step) is the
        optimiser's iteration count."""
        _lib.check(self.lib.sga_bb_step(
            self.ctx, _ptr(x), _ptr(st["y"]), _ptr(st["z_mean"]), _ptr(st["z_logvar"]), _ptr(st["m_y"]), _ptr(st["v_y"]),
            _ptr(st["m_z_mean"]), _ptr(st["v_z_mean"]), _ptr(st["m_z_logvar"]), _ptr(st["v_z_logvar"]), float(temperature),
            int(step), int(step if adam_iterations is None else adam_iterations), _ptr(noise_y), _ptr(eps_z),
            int(bool(rate_only)), float(lr),
            _ptr(self.stats) if want_stats else None, _stream()))

    def eval(self, x, y_tilde, z_mean, z_logvar, step=0, eps_z=None):
        """eval_tensors of bb_sga.py:176-177 at y_tilde = round(y) (bb_sga.py:271-272)."""
        _lib.check(self.lib.sga_bb_eval(self.ctx, _ptr(x), _ptr(y_tilde), _ptr(z_mean), _ptr(z_logvar), int(step),
                                        _ptr(eps_z), _ptr(self.stats), _stream()))
        return self.unpack_stats(self.stats.cpu().numpy())

    def eval_hard(self, *a, **k):
        raise _lib.SgaError("bits-back contexts are evaluated with BbEngine.eval")

    def unpack_stats(self, s):
        per = s[_lib.STATS_HEADER:].reshape(self.n, _lib.BB_STATS_PER_IMAGE)
        return dict(rd_loss=float(s[0]), train_mse=float(s[1]), train_bpp=float(s[2]), psnr=per[:, 0].copy(),
                    est_y_bpp=per[:, 1].copy(), est_z_bpp=per[:, 2].copy(), est_bpp_back=per[:, 5].copy(),
                    est_bpp=per[:, 1] + per[:, 2] - per[:, 5], sse=per[:, 3].copy(), mse=per[:, 4].copy())

    def run(self, x, its=2000, r_its=2000, annealing_rate=1e-3, t0=700, T_ub=0.5, rd_lr=0.005, r_lr=0.003, log_itv=100,
            log=None, log_rate=None, verbose=False):
        """The body of bb_sga.py:200-273 for one batch. Returns (y_tilde, z_mean, z_logvar, eval dict)."""
        # 1. R-D optimisation conditioned on x (bb_sga.py:201-233)
        y = self.encode_init(x)
        zm, zl = self.hyper_init(y)                                   # h_a of the un-rounded y fed as y_tilde (:204)
        st = self.new_state(y, zm, zl)
        for it in range(its):
            T = annealed_temperature(it, r=annealing_rate, ub=T_ub, scheme="exp0", t0=t0)
            at_log = log is not None and (it % log_itv == 0 or it + 1 == its)
            self.step(x, st, T, it, rd_lr, want_stats=at_log)
            if at_log:
                s = self.unpack_stats(self.stats.cpu().numpy())
                # --verbose (bb_sga.py:218-230): the same statistics with y_tilde = round(y) and the current posterior of z
                extra = self.eval(x, torch.round(st["y"]), st["z_mean"], st["z_logvar"], step=it) if verbose else None
                log(it, T, s, extra)
        # 2. y_tilde fixed, rate optimisation w.r.t. z_mean, z_logvar (bb_sga.py:235-256); noise steps continue after
        #    the first loop's so the two loops never share a noise draw
        y_tilde = torch.round(st["y"])                                # np.round: half to even, as torch.round
        zm, zl = self.hyper_init(y_tilde)
        st2 = self.new_state(y_tilde, zm, zl)
        for it in range(r_its):
            at_log = log_rate is not None and (it % log_itv == 0 or it + 1 == r_its)
            self.step(None, st2, 1.0, its + it, r_lr, rate_only=True, want_stats=at_log, adam_iterations=it)
            if at_log:
                log_rate(it, self.unpack_stats(self.stats.cpu().numpy()))
        ev = self.eval(x, y_tilde, st2["z_mean"], st2["z_logvar"], step=its + r_its)
        return y_tilde, st2["z_mean"], st2["z_logvar"], ev
