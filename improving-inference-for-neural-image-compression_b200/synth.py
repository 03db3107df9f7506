"""Seeded synthetic test images (there is no network for Kodak/Tecnick).

iid-uniform pixels are a poor proxy for photographs (latents saturate), so the generator mixes
low-pass-filtered noise, linear gradients and a few hard edges, quantised to uint8."""
import numpy as np


def _smooth_noise(rng, h, w, cells):
    gh, gw = max(2, h // cells + 2), max(2, w // cells + 2)
    g = rng.random((gh, gw, 3))
    yy = np.linspace(0, gh - 1.001, h)
    xx = np.linspace(0, gw - 1.001, w)
    y0, x0 = yy.astype(int), xx.astype(int)
    fy, fx = (yy - y0)[:, None, None], (xx - x0)[None, :, None]
    fy, fx = fy * fy * (3 - 2 * fy), fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def synthetic_images(n, h, w, seed=0):
    """Returns uint8 array [n, h, w, 3]."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, 3), dtype=np.uint8)
    ys = np.linspace(0, 1, h)[:, None, None]
    xs = np.linspace(0, 1, w)[None, :, None]
    for i in range(n):
        img = 0.55 * _smooth_noise(rng, h, w, 64) + 0.25 * _smooth_noise(rng, h, w, 12) + 0.06 * _smooth_noise(rng, h, w, 3)
        gdir = rng.random(3) - 0.5
        img = img + 0.3 * (gdir[0] * ys + gdir[1] * xs)
        for _ in range(3):                                   # a few hard-edged rectangles
            y0, x0 = rng.integers(0, h), rng.integers(0, w)
            hh, ww = rng.integers(h // 8 + 1, h // 2 + 2), rng.integers(w // 8 + 1, w // 2 + 2)
            img[y0:y0 + hh, x0:x0 + ww] += (rng.random(3) - 0.5) * 0.35
        img = img + 0.01 * rng.standard_normal((h, w, 3))
        out[i] = np.clip(np.round(np.clip(img, 0, 1) * 255), 0, 255).astype(np.uint8)
    return out
