"""Model weights in the reference's checkpoint layout.

The reference restores a TF1 checkpoint from ``<checkpoint_dir>/<runname>/`` (sga.py:180-182,
tf_boilerplate.py:47-48) whose variables are created by the Keras layers in nn_models.py.  This
module keeps that directory layout and those variable names; it reads either the reference's own
TF V2 bundle (``checkpoint`` + ``<prefix>.index`` / ``.data-*``, parsed by tf_bundle.py without
TensorFlow) or a ``model.npz`` container (one array per variable).

Variable names (Keras TF1 scopes):
  analysis_transform/layer_{0..3}/{kernel|kernel_rdft, bias}, layer_{0..2}/gdn_{i}/{beta|reparam_beta, gamma|reparam_gamma}
  synthesis_transform/layer_{0..3}/{kernel|kernel_rdft, bias}, layer_{0..2}/igdn_{i}/{...}
  hyper_analysis_transform/layer_{0,1}/{kernel|kernel_rdft, bias}, layer_2/{kernel|kernel_rdft}   (no bias, nn_models.py:95)
  mb_t2018hyper_synthesis_transform/layer_{0..2}/{kernel, bias}                                  (kernel_parameterizer=None, nn_models.py:154)
  entropy_bottleneck/{matrix_0..3, bias_0..3, factor_0..2}
Kernels are (kh, kw, Cin, Cout) like tfc.SignalConv2D; gamma is (in, out).
Re-parameterised variables (``kernel_rdft``, ``reparam_beta``, ``reparam_gamma``) are converted on load.
"""
import os

import numpy as np

GA = "analysis_transform/layer_%d/"
GS = "synthesis_transform/layer_%d/"
HA = "hyper_analysis_transform/layer_%d/"
HS = "mb_t2018hyper_synthesis_transform/layer_%d/"
EB = "entropy_bottleneck/"

GDN_PEDESTAL = 2.0 ** -36          # tfc.NonnegativeParameterizer reparam_offset = 2**-18, pedestal = offset**2
GDN_BETA_MIN = 1e-6
GDN_GAMMA_MIN = 0.0


def variable_shapes(C, hyper_mid=None):
    """name -> shape for the mean-scale hyperprior at `C` filters (sga.py:70-74, nn_models.py)."""
    mid = int(C * 1.5) if hyper_mid is None else hyper_mid       # nn_models.py:157
    s = {}
    for i, (ci, co) in enumerate([(3, C), (C, C), (C, C), (C, C)]):
        s[(GA % i) + "kernel"] = (5, 5, ci, co)
        s[(GA % i) + "bias"] = (co,)
        if i < 3:
            s[(GA % i) + "gdn_%d/beta" % i] = (co,)
            s[(GA % i) + "gdn_%d/gamma" % i] = (co, co)
    for i, (ci, co) in enumerate([(C, C), (C, C), (C, C), (C, 3)]):
        s[(GS % i) + "kernel"] = (5, 5, ci, co)
        s[(GS % i) + "bias"] = (co,)
        if i < 3:
            s[(GS % i) + "igdn_%d/beta" % i] = (co,)
            s[(GS % i) + "igdn_%d/gamma" % i] = (co, co)
    s[(HA % 0) + "kernel"] = (3, 3, C, C)
    s[(HA % 0) + "bias"] = (C,)
    s[(HA % 1) + "kernel"] = (5, 5, C, C)
    s[(HA % 1) + "bias"] = (C,)
    s[(HA % 2) + "kernel"] = (5, 5, C, C)
    s[(HS % 0) + "kernel"] = (5, 5, C, C)
    s[(HS % 0) + "bias"] = (C,)
    s[(HS % 1) + "kernel"] = (5, 5, C, mid)
    s[(HS % 1) + "bias"] = (mid,)
    s[(HS % 2) + "kernel"] = (3, 3, mid, 2 * C)
    s[(HS % 2) + "bias"] = (2 * C,)
    dims = (1, 3, 3, 3, 1)                                        # tfc EntropyBottleneck filters=(3,3,3)
    for i in range(4):
        s[EB + "matrix_%d" % i] = (C, dims[i + 1], dims[i])       # learned_prior.py:45-47
        s[EB + "bias_%d" % i] = (C, dims[i + 1], 1)               # :53-56
        if i < 3:
            s[EB + "factor_%d" % i] = (C, dims[i + 1], 1)         # :60-63
    return s


def make_synthetic(C, seed=0):
    """Random-init weights with realistic statistics (no trained checkpoint is reachable offline):
    variance-preserving (de)conv kernels, beta ~ 1, gamma ~ 0.1 I + small positive, and the
    EntropyBottleneck initialisation of learned_prior.py:36-66 (init_scale=10) plus jitter."""
    rng = np.random.default_rng(seed)
    W = {}
    for name, shp in variable_shapes(C).items():
        leaf = name.split("/")[-1]
        if leaf == "kernel":
            kh, kw, ci, co = shp
            up = name.startswith(GS[:10]) or (name.startswith(HS[:10]) and kh == 5)
            taps = kh * kw / (4.0 if up else 1.0)                  # stride-2 deconv: ~25/4 taps per output
            std = (1.0 / (taps * ci)) ** 0.5
            if name == (GA % 3) + "kernel":
                std *= 6.0                                         # latents y with a spread of a few integers
            if name == (HA % 2) + "kernel":
                std *= 3.0
            if name == (GS % 0) + "kernel":
                std *= 0.55                                        # compensate the spread of y
            if name in ((GS % 1) + "kernel", (GS % 2) + "kernel"):
                std *= 0.75                                        # IGDN expands by ~1.3x per stage
            if name == (GS % 3) + "kernel":
                std *= 0.2                                         # reconstructions with a spread of ~0.2
            W[name] = (rng.standard_normal(shp) * std).astype(np.float32)
        elif leaf == "bias" and not name.startswith(EB):
            b = rng.standard_normal(shp) * 0.05
            if name == (GS % 3) + "bias":
                b = b + 0.45                                       # reconstructions centred in [0,1]
            W[name] = b.astype(np.float32)
        elif leaf == "beta":
            W[name] = (1.0 + 0.1 * rng.random(shp)).astype(np.float32)
        elif leaf == "gamma":
            g = 0.1 * np.eye(shp[0]) + 0.002 * rng.random(shp)
            W[name] = g.astype(np.float32)
    dims = (1, 3, 3, 3, 1)
    scale = 10.0 ** (1 / 4)
    for i in range(4):
        init = np.log(np.expm1(1 / scale / dims[i + 1]))
        W[EB + "matrix_%d" % i] = (init + 0.2 * rng.standard_normal((C, dims[i + 1], dims[i]))).astype(np.float32)
        W[EB + "bias_%d" % i] = rng.uniform(-0.5, 0.5, (C, dims[i + 1], 1)).astype(np.float32)
        if i < 3:
            W[EB + "factor_%d" % i] = (0.3 * rng.standard_normal((C, dims[i + 1], 1))).astype(np.float32)
    return W


# ---- bits-back variant (bb_sga.py / mbt2018_bb.py) ------------------------------------------------------------------
PRIOR = "bmsh_j2018prior/"       # Keras snake-case of learned_prior.BMSHJ2018Prior (bb_sga.py:63)


def variable_shapes_bb(C):
    """bb_sga.py:63-70: h_a emits 2C channels (z_mean | z_logvar, bb_sga.py:99) and the hyperprior density is the
    BMSHJ2018Prior model (same per-channel network as the EntropyBottleneck, learned_prior.py:43-66)."""
    s = {k: v for k, v in variable_shapes(C).items() if not k.startswith(EB)}
    s[(HA % 2) + "kernel"] = (5, 5, C, 2 * C)
    dims = (1, 3, 3, 3, 1)
    for i in range(4):
        s[PRIOR + "matrix_%d" % i] = (C, dims[i + 1], dims[i])
        s[PRIOR + "bias_%d" % i] = (C, dims[i + 1], 1)
        if i < 3:
            s[PRIOR + "factor_%d" % i] = (C, dims[i + 1], 1)
    return s


def make_synthetic_bb(C, seed=0):
    """Synthetic weights of the bits-back model: the mean-scale weights of make_synthetic with the prior renamed and a
    2C-channel last h_a layer whose log-variance half is shifted to give posterior standard deviations around 0.3."""
    W0 = make_synthetic(C, seed)
    rng = np.random.default_rng(seed + 12345)
    W = {}
    for k, v in W0.items():
        if k.startswith(EB):
            W[PRIOR + k[len(EB):]] = v
        elif k == (HA % 2) + "kernel":
            extra = (rng.standard_normal(v.shape) * (0.3 * v.std())).astype(np.float32)
            W[k] = np.concatenate([v, extra], axis=-1)
        else:
            W[k] = v
    return W


# ---- re-parameterisations used by the reference's checkpoints -------------------------------------------

def gdn_from_reparam(reparam, minimum):
    """tfc.NonnegativeParameterizer: var = max(reparam, sqrt(minimum + pedestal))**2 - pedestal."""
    bound = (minimum + GDN_PEDESTAL) ** 0.5
    return (np.maximum(reparam.astype(np.float64), bound) ** 2 - GDN_PEDESTAL).astype(np.float32)


def gdn_to_reparam(var, minimum):
    return np.sqrt(np.maximum(var.astype(np.float64) + GDN_PEDESTAL, minimum + GDN_PEDESTAL)).astype(np.float32)


def irdft_matrix(shape):
    """Orthonormal real-DFT synthesis basis of tfc's RDFTParameterizer (spectral_ops.irdft_matrix, tfc 1.3): a
    (prod(shape), prod(shape)) matrix M with kernel = reshape(M @ kernel_rdft, shape + (Cin, Cout)).

    The basis is SEPARABLE, built one axis at a time: rfft along the axis, the non-DC (and, for even lengths,
    non-Nyquist) bins scaled by sqrt(2), everything divided by sqrt(n_axis), then the real parts of all bins followed by
    the imaginary parts of the scaled bins are concatenated along that same axis (n real coefficients again), and the next
    axis is transformed.  Rows index the flattened spatial position, columns the flattened (per-axis) coefficients.
    Pinned in tests/test_host.py by hand-computed columns; still unchecked against a TF-written checkpoint
    (PARITY_ASSUMPTIONS.md item P9)."""
    shape = tuple(int(s) for s in shape)
    size = int(np.prod(shape))
    rank = len(shape)
    m = np.identity(size, dtype=np.float64).reshape((size,) + shape)
    for axis in range(rank):
        n = shape[axis]
        spec = np.fft.rfft(m, axis=axis + 1)
        sl = [slice(None)] * (rank + 1)
        sl[axis + 1] = slice(1, None) if n % 2 == 1 else slice(1, -1)
        sl = tuple(sl)
        spec[sl] *= np.sqrt(2.0)
        spec /= np.sqrt(n)
        m = np.concatenate([spec.real, spec[sl].imag], axis=axis + 1)
    return m.reshape(size, size)


def deparameterise(raw, shapes=None):
    """Convert any `kernel_rdft` / `reparam_beta` / `reparam_gamma` entries to plain variables.  `shapes` (name -> shape of
    the plain variables) supplies the kernel support when no `kernel_shape` sibling is stored (TF bundles)."""
    W = {}
    for name, arr in raw.items():
        arr = np.asarray(arr)
        head, leaf = name.rsplit("/", 1) if "/" in name else ("", name)
        if leaf == "reparam_beta":
            W[head + "/beta"] = gdn_from_reparam(arr, GDN_BETA_MIN)
        elif leaf == "reparam_gamma":
            W[head + "/gamma"] = gdn_from_reparam(arr, GDN_GAMMA_MIN)
        elif leaf == "kernel_rdft":
            shp = raw[head + "/kernel_shape"] if head + "/kernel_shape" in raw else None
            if shp is None and shapes is not None:
                shp = shapes.get(head + "/kernel")
            if shp is None:
                raise ValueError("kernel_rdft needs a sibling kernel_shape entry: %s" % name)
            kh, kw, ci, co = (int(v) for v in shp)
            W[head + "/kernel"] = (irdft_matrix((kh, kw)) @ arr.astype(np.float64)).reshape(kh, kw, ci, co).astype(np.float32)
        elif leaf == "kernel_shape":
            continue
        else:
            W[name] = arr.astype(np.float32)
    return W


def check_complete(W, C, bb=False):
    want = variable_shapes_bb(C) if bb else variable_shapes(C)
    missing = [k for k in want if k not in W]
    if missing:
        raise KeyError("checkpoint is missing variables: %s" % ", ".join(missing[:6]))
    for k, shp in want.items():
        if tuple(W[k].shape) != tuple(shp):
            raise ValueError("variable %s has shape %s, expected %s (num_filters=%d)" % (k, W[k].shape, shp, C))


def save_checkpoint(checkpoint_dir, runname, W):
    d = os.path.join(checkpoint_dir, runname)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "model.npz")
    np.savez(path, **W)
    with open(os.path.join(d, "checkpoint"), "w") as f:          # same marker file TF writes
        f.write('model_checkpoint_path: "model.npz"\n')
    return path


def load_checkpoint(checkpoint_dir, runname, num_filters, bb=False):
    """Mirror of tf.train.latest_checkpoint + Saver.restore (sga.py:180-182)."""
    d = os.path.join(checkpoint_dir, runname)
    path = os.path.join(d, "model.npz")
    want = variable_shapes_bb(num_filters) if bb else variable_shapes(num_filters)
    if os.path.exists(path):
        with np.load(path) as f:
            raw = {k: f[k] for k in f.files}
    else:
        # the reference's own format: a TF V2 bundle named by the `checkpoint` state file (tf_bundle.py)
        from . import tf_bundle
        prefix = tf_bundle.latest_checkpoint(d)
        if prefix is None or not os.path.exists(prefix + ".index"):
            raise FileNotFoundError("no model.npz and no TensorFlow checkpoint (checkpoint + <prefix>.index/.data-*) under %s" % d)
        heads = {k.rsplit("/", 1)[0] for k in want}
        raw = tf_bundle.read_bundle(prefix, names=lambda n: n.rsplit("/", 1)[0] in heads and
                                    n.rsplit("/", 1)[-1] in ("kernel", "kernel_rdft", "bias", "beta", "gamma", "reparam_beta",
                                                             "reparam_gamma") or n in want)
    W = deparameterise(raw, shapes=want)
    check_complete(W, num_filters, bb=bb)
    return W


def runname(num_filters, lmbda, prefix="mbt2018"):
    """utils.get_runname format (utils.py:51-69) for record_keys (num_filters, lmbda)."""
    return "%s-num_filters=%s-lmbda=%s" % (prefix, num_filters, lmbda)
