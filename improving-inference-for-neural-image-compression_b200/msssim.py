"""MS-SSIM of the final reconstruction, once per batch (not on the hot path).

Restates tf.image.ssim_multiscale as called at sga.py:175-176: x_hat = round(clip(x_tilde,0,1)*255) vs 255*x,
max_val 255, 5 scales with power factors (0.0448, 0.2856, 0.3001, 0.2363, 0.1333), 11x11 Gaussian (sigma 1.5), k1=0.01,
k2=0.03, 2x2 average pooling between scales, per-channel product then mean over channels.  TF-behaviour recalled from
TF 1.15's image_ops_impl.py (PARITY_ASSUMPTIONS.md P8).  Runs with torch ops on whichever device holds the tensors."""
import numpy as np
import torch
import torch.nn.functional as F

POWER_FACTORS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _gauss_kernel(size, sigma, dtype, device):
    coords = torch.arange(size, dtype=dtype, device=device) - (size - 1) / 2.0
    g = -0.5 * coords ** 2 / sigma ** 2
    g = g.view(1, -1) + g.view(-1, 1)
    return torch.softmax(g.view(-1), 0).view(1, 1, size, size)


def _ssim_cs(a, b, max_val, size=11, sigma=1.5, k1=0.01, k2=0.03):
    """a, b: [N,C,H,W]. Returns (ssim, cs) per image and channel: [N,C]."""
    c = a.shape[1]
    k = _gauss_kernel(size, sigma, a.dtype, a.device).repeat(c, 1, 1, 1)
    red = lambda t: F.conv2d(t, k, groups=c)
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    m0, m1 = red(a), red(b)
    num0, den0 = m0 * m1 * 2.0, m0 * m0 + m1 * m1
    lum = (num0 + c1) / (den0 + c1)
    num1, den1 = red(a * b) * 2.0, red(a * a + b * b)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).mean(dim=(2, 3)), cs.mean(dim=(2, 3))


def msssim_per_image(x_tilde, x, max_val=255.0):
    """x_tilde, x: [N,H,W,3] float (x in [0,1]).  Returns numpy [N]; NaN if the image is too small for 5 scales."""
    a = torch.round(torch.clamp(x_tilde, 0, 1) * 255).permute(0, 3, 1, 2).double()
    b = (x * 255).permute(0, 3, 1, 2).double()
    n = a.shape[0]
    if min(a.shape[2], a.shape[3]) < 11 * 2 ** (len(POWER_FACTORS) - 1):
        return np.full((n,), np.nan, dtype=np.float32)
    mcs = []
    for i in range(len(POWER_FACTORS)):
        if i > 0:
            ph, pw = a.shape[2] % 2, a.shape[3] % 2
            if ph or pw:                                   # TF pads odd sizes symmetrically before pooling
                a = F.pad(a, (0, pw, 0, ph), mode="replicate")
                b = F.pad(b, (0, pw, 0, ph), mode="replicate")
            a, b = F.avg_pool2d(a, 2), F.avg_pool2d(b, 2)
        ssim, cs = _ssim_cs(a, b, max_val)
        mcs.append(torch.relu(cs))
    mcs[-1] = torch.relu(ssim)
    stack = torch.stack(mcs, dim=-1)                        # [N,C,5]
    pw = torch.tensor(POWER_FACTORS, dtype=stack.dtype, device=stack.device)
    ms = torch.prod(stack ** pw, dim=-1).mean(dim=1)
    return ms.float().cpu().numpy()
