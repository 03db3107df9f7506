"""CPU tests of the oracle itself: known-answer vectors, the reference's own importable pieces, autograd checks of
every hand-derived backward, independent re-implementations of the index conventions, and invariants."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import philox, sga_oracle as O
from helpers import make_case, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32_10."""
    for ctr, key, want in philox.KAT:
        got = philox.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_philox_uniform_range_and_layout():
    u0, u1 = philox.sga_uniforms(1001, seed=42, step=5, tensor_id=1)
    assert u0.shape == (1001,) and u0.dtype == np.float32
    assert u0.min() > 0 and u0.max() < 1 and u1.min() > 0 and u1.max() < 1
    v0, _ = philox.sga_uniforms(1001, seed=42, step=6, tensor_id=1)
    assert not np.array_equal(u0, v0)
    a0, _ = philox.sga_uniforms(500, seed=42, step=5, tensor_id=1)      # counter-based: prefix property
    assert np.array_equal(a0, u0[:500])
    assert abs(u0.mean() - 0.5) < 0.05


def test_temperature_schedule_matches_reference_values():
    """utils.annealed_temperature (utils.py:166-180) with the sga.py:190-196 settings; values quoted in SURVEY a1."""
    f = lambda t: O.annealed_temperature(t, r=1e-3, ub=0.5, scheme="exp0", t0=700)
    assert f(0) == 0.5 and f(700) == 0.5
    assert abs(f(1000) - 0.5 * math.exp(-0.3)) < 1e-15
    assert abs(f(1999) - 0.1364) < 1e-4
    assert O.annealed_temperature(10 ** 6, r=1.0, ub=0.5, scheme="exp0", t0=0) == 1e-8
    from sga_b200.schedule import annealed_temperature as host_sched
    for t in (0, 1, 699, 701, 1234, 1999):
        assert host_sched(t, r=1e-3, ub=0.5, scheme="exp0", t0=700) == f(t)


def test_adam_matches_reference_golden():
    """oracle Adam == the reference's adam.py (fixture written by tests/golden/make_golden.py)."""
    gold = np.load(os.path.join(GOLD, "adam_ref.npz"))
    opt = O.Adam(lr=0.005)
    p = [torch.from_numpy(gold["p0"].copy())]
    for it in range(gold["grads"].shape[0]):
        p = opt.update(p, [torch.from_numpy(gold["grads"][it])])
        assert p[0].dtype == torch.float32
        ref = gold["params"][it]
        assert np.abs(p[0].numpy() - ref).max() <= np.spacing(np.abs(ref).max().astype(np.float32))


@pytest.mark.skipif(not os.path.exists("/root/reference/adam.py"), reason="reference checkout not present")
def test_adam_against_live_reference():
    from golden.make_golden import reference_adam_class
    Adam = reference_adam_class()
    rng = np.random.default_rng(3)
    p0 = rng.standard_normal(777).astype(np.float32)
    ref, mine = Adam(lr=0.005), O.Adam(lr=0.005)
    pr, pm = [p0.copy()], [torch.from_numpy(p0.copy())]
    for _ in range(5):
        g = rng.standard_normal(777).astype(np.float32)
        pr = ref.update(pr, [g])
        pm = mine.update(pm, [torch.from_numpy(g)])
    assert np.abs(pr[0] - pm[0].numpy()).max() <= np.spacing(np.float32(np.abs(pr[0]).max()))


def test_golden_sga_fixture_reproduces():
    """The committed oracle outputs are reproduced by the current oracle (guards against silent edits)."""
    import sga_b200
    gold = np.load(os.path.join(GOLD, "sga_c16.npz"))
    C, n, h, w, its, seed = (int(gold[k]) for k in ("C", "n", "h", "w", "its", "seed"))
    W = sga_b200.weights.make_synthetic(C, seed)
    x = torch.from_numpy(sga_b200.synth.synthetic_images(n, h, w, seed).astype(np.float32) / 255.0)
    rec = []
    y, z, ev = O.sga_run(x, W, float(gold["lmbda"]), its, seed=seed, annealing_rate=float(gold["annealing_rate"]),
                         t0=int(gold["t0"]), record=rec)
    assert np.allclose([float(r["rd_loss"]) for r in rec], gold["rd_loss"], rtol=1e-5)
    assert rel_err(rec[0]["grad_y"], gold["grad_y0"]) < 1e-4
    assert np.abs(ev["psnr"].numpy() - gold["eval_psnr"]).max() < 1e-3
    assert np.abs(ev["est_bpp"].numpy() - gold["eval_est_bpp"]).max() < 1e-4


@pytest.mark.parametrize("T,lmbda", [(0.5, 0.01), (0.15, 0.08), (0.3, 0.0)])
def test_hand_derived_backward_equals_autograd(T, lmbda):
    """tf.gradients(rd_loss, [y, z]) (sga.py:164): every hand-written backward vs torch.autograd in float64."""
    W, _, x, y, z = make_case(8, 2, 32, 48, seed=1)
    Wd = O._to(W, torch.float64)
    x, y, z = x.double(), y.double(), z.double()
    noise = O.draw_noise(y, z, 0, 3)
    a = O.sga_graph(x, y, z, Wd, T, noise, lmbda)
    b = O.sga_graph_autograd(x, y, z, Wd, T, noise, lmbda)
    assert abs(float(a["rd_loss"]) - float(b["rd_loss"])) < 1e-9 * abs(float(b["rd_loss"]))
    assert rel_err(a["grad_y"], b["grad_y"]) < 1e-9 and rel_err(a["grad_z"], b["grad_z"]) < 1e-9


def test_crop_path_backward_equals_autograd():
    """75x50 image: y is 5x4, z 2x1, h_s output 8x4 cropped to 5x4 (the sga.py:126-128 crop)."""
    W, _, x, y, z = make_case(8, 1, 75, 50, seed=2)
    Wd = O._to(W, torch.float64)
    x, y, z = x.double(), y.double(), z.double()
    noise = O.draw_noise(y, z, 1, 0)
    a = O.sga_graph(x, y, z, Wd, 0.4, noise, 0.02)
    b = O.sga_graph_autograd(x, y, z, Wd, 0.4, noise, 0.02)
    assert rel_err(a["grad_y"], b["grad_y"]) < 1e-9 and rel_err(a["grad_z"], b["grad_z"]) < 1e-9


def _deconv_direct(x, k):
    """out[t] = sum_i x[i] k[t + 2 - 2i] written as explicit loops (independent of torch's conv)."""
    n, h, w, ci = x.shape
    kh, kw, _, co = k.shape
    out = torch.zeros(n, 2 * h, 2 * w, co, dtype=x.dtype)
    for ty in range(2 * h):
        for tx in range(2 * w):
            for i in range(h):
                a = ty + kh // 2 - 2 * i
                if not 0 <= a < kh:
                    continue
                for j in range(w):
                    b = tx + kw // 2 - 2 * j
                    if 0 <= b < kw:
                        out[:, ty, tx, :] += x[:, i, j, :] @ k[a, b]
    return out


def test_deconv_index_convention_and_adjoint():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 4, 5, generator=g, dtype=torch.float64)
    k = torch.randn(5, 5, 5, 6, generator=g, dtype=torch.float64)
    out = O.deconv_s2(x, k)
    assert tuple(out.shape) == (1, 6, 8, 6)                      # exactly 2x
    assert rel_err(out, _deconv_direct(x, k)) < 1e-12
    gy = torch.randn(out.shape, generator=g, dtype=torch.float64)
    assert abs((out * gy).sum() - (x * O.deconv_s2_dgrad(gy, k)).sum()) < 1e-9      # <Ax, y> == <x, A^T y>
    k3 = torch.randn(3, 3, 5, 6, generator=g, dtype=torch.float64)
    o3 = O.conv_s1_true(x, k3)
    g3 = torch.randn(o3.shape, generator=g, dtype=torch.float64)
    assert abs((o3 * g3).sum() - (x * O.conv_s1_true_dgrad(g3, k3)).sum()) < 1e-9
    # true convolution == correlation with the flipped kernel
    assert rel_err(o3, O.corr_down(x, torch.flip(k3, dims=(0, 1)), None, 1)) < 1e-12


def test_igdn_backward_and_entropy_model_invariants():
    g = torch.Generator().manual_seed(1)
    C = 6
    x = torch.randn(2, 3, 3, C, generator=g, dtype=torch.float64, requires_grad=True)
    beta = 1 + torch.rand(C, generator=g, dtype=torch.float64)
    gamma = 0.1 * torch.eye(C, dtype=torch.float64) + 0.05 * torch.rand(C, C, generator=g, dtype=torch.float64)
    out, s = O.igdn(x, beta, gamma)
    gup = torch.randn(out.shape, generator=g, dtype=torch.float64)
    (auto,) = torch.autograd.grad((out * gup).sum(), x)
    assert rel_err(O.igdn_bwd(gup, x.detach(), s.detach(), gamma), auto) < 1e-12
    # both entropy models are probability mass functions over the integers
    import sga_b200
    W = sga_b200.weights.make_synthetic(8, 0)
    eb = {k: torch.from_numpy(v).double() for k, v in O.eb_params(W).items()}
    ks = torch.arange(-80, 81, dtype=torch.float64).view(-1, 1).repeat(1, 8)
    p, _ = O.eb_likelihood(ks, eb)
    # telescoping: sum_k [cdf(k+.5) - cdf(k-.5)] == cdf(K+.5) - cdf(-K-.5); the tails are logistic (heavy)
    hi, _ = O.eb_logits_and_slope(torch.full((1, 8), 80.5, dtype=torch.float64), eb, want_slope=False)
    lo, _ = O.eb_logits_and_slope(torch.full((1, 8), -80.5, dtype=torch.float64), eb, want_slope=False)
    assert (p.sum(0) - (torch.sigmoid(hi) - torch.sigmoid(lo))[0]).abs().max() < 1e-12
    assert (p.sum(0) - 1).abs().max() < 1e-2
    ks1 = torch.arange(-300, 301, dtype=torch.float64)
    for mu, ls in [(0.3, -3.0), (-2.2, 0.7), (5.0, 3.0)]:
        pg = O.gauss_likelihood(ks1, torch.full_like(ks1, mu), torch.full_like(ks1, ls))[0]
        assert abs(pg.sum() - 1) < 1e-9


def test_soft_round_limits():
    """T -> small: soft round approaches hard rounding decisions; integers pass through with zero gradient."""
    v = torch.tensor([0.2, 0.8, -1.3, 2.0, -4.0, 0.5000001], dtype=torch.float64)
    u = torch.full_like(v, 0.5)
    vt, dv = O.soft_round(v, u, u, 0.02)
    assert torch.allclose(vt[:3], torch.tensor([0.0, 1.0, -1.0], dtype=torch.float64), atol=1e-6)
    assert vt[3] == 2.0 and vt[4] == -4.0 and dv[3] == 0 and dv[4] == 0
    vt2, dv2 = O.soft_round(v, u, u, 0.5)
    auto = torch.autograd.functional.jacobian(lambda t: O.soft_round_autograd(t, u, u, 0.5), v).diagonal()
    assert rel_err(dv2, auto) < 1e-10
