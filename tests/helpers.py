"""Shared test helpers: tiny models, oracle <-> device conversion, error metrics."""
import numpy as np
import torch

import sga_b200
from sga_b200 import synth, weights
from oracle import sga_oracle as O


def rel_err(a, b):
    """max |a-b| / (max |b| + tiny): scale-relative max error."""
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def rms_rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30)).item()


def make_case(C, n, h, w, seed=0, jitter=0.3):
    """Synthetic weights + images + perturbed latents (so that y is not sitting at g_a(x))."""
    W = weights.make_synthetic(C, seed)
    x = torch.from_numpy(synth.synthetic_images(n, h, w, seed).astype(np.float32) / 255.0)
    Wt = O._to(W, torch.float32)
    y, z = O.analysis(x, Wt)
    g = torch.Generator().manual_seed(seed + 1)
    y = y + jitter * torch.randn(y.shape, generator=g)
    z = z + jitter * torch.randn(z.shape, generator=g)
    return W, Wt, x, y.contiguous(), z.contiguous()
