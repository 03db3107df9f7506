"""Parity of the element-wise CUDA kernels (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import philox, sga_oracle as O
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    from sga_b200 import ops
    return ops


@pytest.mark.parametrize("numel,T", [(1, 0.5), (7, 0.5), (4096, 0.5), (100003, 0.1364), (65536, 0.05)])
def test_soft_round_philox(numel, T):
    """a2 (sga.py:88-98): same Philox stream on device and in the oracle; ragged and odd sizes."""
    ops = _ops()
    g = torch.Generator().manual_seed(numel)
    v = (torch.randn(numel, generator=g) * 3).float()
    v[::17] = torch.round(v[::17])            # exact integers: floor == ceil, zero gradient
    u0, u1 = philox.sga_uniforms(numel, 1234567890123, 7, 2)
    vt_o, dv_o = O.soft_round(v, torch.from_numpy(u0), torch.from_numpy(u1), T)
    vt, dv = ops.soft_round(v.cuda(), T, seed=1234567890123, step=7, tensor_id=2)
    assert torch.isfinite(vt).all() and torch.isfinite(dv).all()
    # float32 transcendental chains (atanh/log/exp) differ by a few ulp between libm and CUDA
    assert (vt.cpu() - vt_o).abs().max().item() < 2e-5
    assert rel_err(dv, dv_o) < 2e-4
    ints = v == torch.round(v)     # floor == ceil: v_tilde = f*s0 + f*s1 == v up to one rounding, gradient exactly 0
    assert (vt.cpu()[ints] - v[ints]).abs().max().item() <= 1e-6 * v.abs().max().item()
    assert (dv.cpu()[ints] == 0).all()


def test_soft_round_injected_noise():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    v = (torch.randn(5000, generator=g) * 2).float()
    u = torch.rand(2, 5000, generator=g).clamp(1e-6, 1 - 1e-6)
    vt_o, dv_o = O.soft_round(v, u[0], u[1], 0.3)
    vt, dv = ops.soft_round(v.cuda(), 0.3, u=u.contiguous().cuda())
    assert (vt.cpu() - vt_o).abs().max().item() < 2e-5
    assert rel_err(dv, dv_o) < 2e-4


@pytest.mark.parametrize("C", [8, 192])
def test_eb_likelihood(C):
    """a3: factorised prior (tfc EntropyBottleneck._likelihood; learned_prior.py:78-121, 296-347)."""
    import sga_b200
    ops = _ops()
    W = sga_b200.weights.make_synthetic(C, 3)
    eb = {k.split("/")[-1]: torch.from_numpy(v) for k, v in W.items() if k.startswith("entropy_bottleneck/")}
    g = torch.Generator().manual_seed(1)
    zt = (torch.randn(3, 5, 4, C, generator=g) * 4).float()
    p_o, dp_o = O.eb_likelihood(zt.double(), {k: v.double() for k, v in eb.items()})
    p, dp = ops.eb_likelihood(zt.cuda(), [eb["matrix_%d" % i].cuda() for i in range(4)],
                              [eb["bias_%d" % i].cuda() for i in range(4)], [eb["factor_%d" % i].cuda() for i in range(3)])
    assert rel_err(p, p_o) < 1e-5
    assert rel_err(dp, dp_o) < 1e-4
    # the model is a proper pmf over the integers (size-independent property)
    # (logistic tails with a scale of ~10: +-3000 covers all but ~1e-4 of the mass)
    ks = torch.arange(-3000, 3001).float().view(-1, 1).repeat(1, C)
    pk, _ = ops.eb_likelihood(ks.cuda().contiguous(), [eb["matrix_%d" % i].cuda() for i in range(4)],
                              [eb["bias_%d" % i].cuda() for i in range(4)], [eb["factor_%d" % i].cuda() for i in range(3)])
    pk_o, _ = O.eb_likelihood(ks.double(), {k: v.double() for k, v in eb.items()})
    assert (pk.double().sum(0).cpu() - pk_o.sum(0)).abs().max().item() < 1e-4
    assert (pk.double().sum(0).cpu() - 1).abs().max().item() < 2e-3


def test_gauss_likelihood():
    """a6: box-convolved Gaussian (utils.py:80-102) incl. the 0.11 scale bound and far tails."""
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    n = 20000
    yt = torch.round(torch.randn(n, generator=g) * 5).float()
    mu = (torch.randn(n, generator=g) * 2).float()
    ls = (torch.randn(n, generator=g) * 1.5 - 0.5).float()       # sigma from ~0.02 to ~30
    yt[:10] = mu[:10]                                            # |y-mu| = 0 -> sign 0
    p_o, dy_o, _, ds_o, _, _ = O.gauss_likelihood(yt.double(), mu.double(), ls.double())
    p, dy, ds = ops.gauss_likelihood(yt.cuda(), mu.cuda(), ls.cuda())
    big = p_o > 1e-6
    assert ((p.cpu().double() - p_o).abs()[big] / p_o[big]).max().item() < 2e-4
    assert (p.cpu().double() - p_o).abs().max().item() < 1e-6
    assert rel_err(dy, dy_o) < 1e-4 and rel_err(ds, ds_o) < 1e-4
    # pmf over the integers for fixed (mu, sigma)
    ks = torch.arange(-200, 201).float()
    for m, s in [(0.3, -2.0), (-1.7, 0.5), (4.2, 2.5)]:
        pk, _, _ = ops.gauss_likelihood(ks.cuda(), torch.full_like(ks, m).cuda(), torch.full_like(ks, s).cuda())
        assert abs(pk.sum().item() - 1) < 1e-4


def test_adam_matches_reference_golden():
    """a12: bit-level agreement with the reference's own adam.py (float32 state) -- tests/golden/adam_ref.npz."""
    import os
    ops = _ops()
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "adam_ref.npz"))
    p = torch.from_numpy(gold["p0"]).cuda()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for it in range(gold["grads"].shape[0]):
        ops.adam_(p, m, v, torch.from_numpy(gold["grads"][it]).cuda(), 0.005, 0.9, 0.999, 1e-8, it)
        ref = gold["params"][it]
        diff = np.abs(p.cpu().numpy() - ref)
        # identical operation order and hyper-parameters carried as doubles (float32(1 - beta) formed as numpy forms it):
        # at most 1 ulp PER ELEMENT (a last-bit difference of the host pow() in lr_t), exact for nearly all elements
        assert (diff <= np.spacing(np.abs(ref).astype(np.float32))).all(), (it, diff.max())
        assert (diff == 0).mean() > 0.99, (it, (diff == 0).mean())


def test_adam_empty():
    ops = _ops()
    e = torch.empty(0, device="cuda")
    ops.adam_(e, e.clone(), e.clone(), e.clone(), 0.005, 0.9, 0.999, 1e-8, 0)
