"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no compute without a GPU),
weights container, CLI surface, batch segmentation / sharding logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from sga_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "sga_b200.h")).read()
    declared = set(re.findall(r"\b(sga_[a-z0-9_]+)\s*\(", header))
    declared -= {"sga_status"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sga_version() >= 100
    assert ctypes.sizeof(_lib.SgaConfig) == 80


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from sga_b200 import SgaError, weights
    from sga_b200.engine import SgaEngine
    with pytest.raises(SgaError):
        SgaEngine(weights.make_synthetic(8, 0), 1, 32, 32, 8, 0.01)


def test_argument_errors_are_reported_not_thrown():
    from sga_b200 import _lib
    lib = _lib.load()
    cfg = _lib.SgaConfig(n_images=0, height=16, width=16, num_filters=8, n_ref=1, operand=0)
    ctx = ctypes.c_void_p()
    assert lib.sga_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1
    assert b"shape" in lib.sga_last_error()
    cfg = _lib.SgaConfig(n_images=1, height=16, width=16, num_filters=100, n_ref=1, operand=1)
    assert lib.sga_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1       # bf16 path needs C % 64 == 0
    assert lib.sga_destroy(None) == 0


def test_weights_layout_and_roundtrip(tmp_path):
    from sga_b200 import weights
    W = weights.make_synthetic(16, 1)
    shapes = weights.variable_shapes(16)
    assert set(W) == set(shapes)
    assert shapes["mb_t2018hyper_synthesis_transform/layer_1/kernel"] == (5, 5, 16, 24)      # int(1.5 C)
    assert shapes["mb_t2018hyper_synthesis_transform/layer_2/kernel"] == (3, 3, 24, 32)
    assert "hyper_analysis_transform/layer_2/bias" not in shapes                             # nn_models.py:95
    assert shapes["entropy_bottleneck/matrix_1"] == (16, 3, 3)
    rn = weights.runname(16, 0.01)
    assert rn == "mbt2018-num_filters=16-lmbda=0.01"
    weights.save_checkpoint(str(tmp_path), rn, W)
    W2 = weights.load_checkpoint(str(tmp_path), rn, 16)
    assert all(np.array_equal(W[k], W2[k]) for k in W)
    with pytest.raises(ValueError):
        weights.load_checkpoint(str(tmp_path), rn, 32)
    with pytest.raises(FileNotFoundError):
        weights.load_checkpoint(str(tmp_path), "nope", 16)


def test_reparameterised_checkpoint_entries():
    from sga_b200 import weights
    rng = np.random.default_rng(0)
    beta = (1 + rng.random(8)).astype(np.float32)
    gamma = (0.1 * rng.random((8, 8))).astype(np.float32)
    raw = {"a/gdn_0/reparam_beta": weights.gdn_to_reparam(beta, weights.GDN_BETA_MIN),
           "a/gdn_0/reparam_gamma": weights.gdn_to_reparam(gamma, weights.GDN_GAMMA_MIN)}
    W = weights.deparameterise(raw)
    assert np.allclose(W["a/gdn_0/beta"], beta, rtol=1e-6) and np.allclose(W["a/gdn_0/gamma"], gamma, atol=1e-7)
    M = weights.irdft_matrix((5, 5))
    assert M.shape == (25, 25) and np.allclose(M.T @ M, np.eye(25), atol=1e-9)     # orthonormal basis
    # hand-built 1-D basis of tfc's per-axis construction: [real bins 0..n/2 | imaginary parts of the sqrt(2)-scaled bins]
    def basis_1d(n):
        k = np.arange(n)[:, None]
        nb = n // 2 + 1
        scaled = list(range(1, nb)) if n % 2 else list(range(1, nb - 1))
        cols = []
        for b in range(nb):
            cols.append(np.cos(2 * np.pi * k * b / n) * (np.sqrt(2.0) if b in scaled else 1.0))
        for b in scaled:
            cols.append(-np.sin(2 * np.pi * k * b / n) * np.sqrt(2.0))
        return np.concatenate(cols, axis=1) / np.sqrt(n)
    assert np.allclose(weights.irdft_matrix((3,)), [[0.57735027, 0.81649658, 0.0], [0.57735027, -0.40824829, -0.70710678],
                                                    [0.57735027, -0.40824829, 0.70710678]], atol=1e-7)
    for n in (2, 3, 4, 5):
        assert np.allclose(weights.irdft_matrix((n,)), basis_1d(n), atol=1e-12)
    # 2-D: separable -- row (i, j), column (a, b) = B_h[i, a] * B_w[j, b]
    assert np.allclose(M, np.kron(basis_1d(5), basis_1d(5)), atol=1e-12)
    assert np.allclose(weights.irdft_matrix((3, 4)), np.kron(basis_1d(3), basis_1d(4)), atol=1e-12)
    k = rng.standard_normal((5, 5, 2, 3)).astype(np.float32)
    rdft = M.T @ k.reshape(25, 6)
    W = weights.deparameterise({"l/kernel_rdft": rdft, "l/kernel_shape": np.array([5, 5, 2, 3])})
    assert np.allclose(W["l/kernel"], k, atol=1e-5)


def test_cli_surface_matches_reference():
    """tf_boilerplate.parse_args (tf_boilerplate.py:91-204): same flags, dests and defaults for `compress`."""
    from sga_b200.cli import parse_args
    a = parse_args(["sga.py", "--num_filters", "192", "compress", "mbt2018-num_filters=192-lmbda=0.01", "in.npy"])
    assert a.command == "compress" and a.num_filters == 192 and a.lmbda == -1 and a.sga_its == 2000
    assert a.annealing_rate == 1e-3 and a.t0 == 700 and a.checkpoint_dir == "./checkpoints" and a.results_dir == "./results"
    assert a.verbose is False and a.output_file is None and a.num_hfilters == -1
    a = parse_args(["sga.py", "-V", "--num_filters", "256", "--checkpoint_dir", "/c", "compress", "--lambda", "0.08",
                    "--sga_its", "10", "--annealing_rate", "2e-3", "--t0", "350", "--results_dir", "/r", "run", "x.png", "o"])
    assert a.verbose and a.lmbda == 0.08 and a.sga_its == 10 and a.t0 == 350 and a.output_file == "o"
    with pytest.raises(SystemExit):
        parse_args(["sga.py"])
    assert a.save_opt_record is False and a.save_reconstruction is False      # configs.save_opt_record / sga.py:34 as flags
    a = parse_args(["sga.py", "--num_filters", "192", "compress", "--save_opt_record", "--save_reconstruction", "run", "x.png"])
    assert a.save_opt_record and a.save_reconstruction
    from sga_b200.cli import parse_args_mbt2018
    m = parse_args_mbt2018(["mbt2018.py", "--num_filters", "192", "compress", "run", "in.png", "out.tfci"])
    assert m.command == "compress" and m.output_file == "out.tfci" and m.results_dir == "./results"
    m = parse_args_mbt2018(["mbt2018.py", "--num_filters", "192", "decompress", "run", "in.tfci"])
    assert m.command == "decompress" and m.output_file is None
    from sga_b200 import configs
    assert configs.get_eval_batch_size(768 * 512) == 25 and configs.get_eval_batch_size(1200 * 1200) == 7   # configs.py:8-9
    assert configs.save_opt_record is False and configs.write_tfci_for_eval is False
    b = parse_args(["bb_sga.py", "--num_filters", "192", "compress", "mbt2018_bb-num_filters=192-lmbda=0.01", "in.npy"], bb=True)
    assert b.sga_its == 2000 and b.rate_its == 2000 and b.lmbda == -1          # bb_sga.py:186-192


def test_bits_back_weights_layout(tmp_path):
    """bb_sga.py:63-70: 2C-channel h_a output, BMSHJ2018Prior variables instead of the EntropyBottleneck's."""
    from sga_b200 import weights as Wm
    W = Wm.make_synthetic_bb(8, 0)
    assert W["hyper_analysis_transform/layer_2/kernel"].shape == (5, 5, 8, 16)
    assert "bmsh_j2018prior/matrix_0" in W and not any(k.startswith("entropy_bottleneck/") for k in W)
    rn = Wm.runname(8, 0.01, prefix="mbt2018_bb")
    Wm.save_checkpoint(str(tmp_path), rn, W)
    W2 = Wm.load_checkpoint(str(tmp_path), rn, 8, bb=True)
    assert set(W2) == set(W) and all(np.array_equal(W2[k], W[k]) for k in W)
    with pytest.raises(KeyError):
        Wm.load_checkpoint(str(tmp_path), rn, 8)          # not a mean-scale (sga.py) checkpoint


def test_batch_segmentation_and_shards():
    """configs.get_eval_batch_size (configs.py:8-9) and the contiguous sharding of each reference batch."""
    from sga_b200.dist import eval_batch_size, reference_batches, shard_range
    assert eval_batch_size(768 * 512) == 25 and eval_batch_size(1200 * 1200) == 7 and eval_batch_size(256 * 256) == 153
    assert reference_batches(24, 768 * 512) == [(0, 24)]
    b = reference_batches(100, 1200 * 1200)
    assert b[0] == (0, 7) and b[-1] == (98, 100) and len(b) == 15
    for n, ws in [(24, 8), (7, 8), (2, 8), (25, 4), (1, 1)]:
        cover = []
        for r in range(ws):
            lo, hi = shard_range(n, r, ws)
            assert 0 <= lo <= hi <= n
            cover += list(range(lo, hi))
        assert cover == list(range(n))
    assert [shard_range(24, r, 8) for r in range(8)] == [(3 * r, 3 * r + 3) for r in range(8)]


def test_gather_results_gloo_world2(tmp_path):
    """The only collective of the path (one gather of per-image rows after the loop) on CPU with gloo, world_size 2."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        "sys.path.insert(0, %r)\n"
        "from sga_b200.dist import gather_rows, shard_range, allreduce_scalars, allgather_floats, rank_segments\n"
        "dist.init_process_group('gloo')\n"
        "r, ws = dist.get_rank(), dist.get_world_size()\n"
        "n = 5\n"
        "lo, hi = shard_range(n, r, ws)\n"
        "segs = rank_segments(n, 1200 * 1200 * 4, r, ws)           # batches of 2 images: 2 + 2 + 1\n"
        "assert sum(s['hi'] - s['lo'] for s in segs) == hi - lo and segs[0]['lo'] == lo\n"
        "per_batch = [0.0, 0.0, 0.0]\n"
        "for s in segs: per_batch[s['batch']] += s['hi'] - s['lo']\n"
        "assert allreduce_scalars(per_batch, device='cpu') == [2.0, 2.0, 1.0]   # the per-batch log reduction of cli.compress\n"
        "assert allgather_floats(float(r + 1)) == [1.0, 2.0]\n"
        "rows = np.stack([np.arange(lo, hi) * 10.0 + k for k in range(5)], axis=1).astype(np.float32)\n"
        "full = gather_rows(rows, n, r, ws, device='cpu')\n"
        "tot = allreduce_scalars([float(hi - lo), 1.0], device='cpu')\n"
        "if r == 0:\n"
        "    want = np.stack([np.arange(n) * 10.0 + k for k in range(5)], axis=1)\n"
        "    assert np.array_equal(full, want), full\n"
        "    assert tot == [5.0, 2.0], tot\n"
        "    print('GATHER_OK')\n"
        "dist.destroy_process_group()\n" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert "GATHER_OK" in r.stdout, r.stdout + r.stderr


def test_tf_bundle_roundtrip_and_checkpoint_loading(tmp_path):
    """sga.py:180-182: <checkpoint_dir>/<runname>/{checkpoint, <prefix>.index, <prefix>.data-00000-of-00001}.  The
    writer/reader pair restates the TF V2 bundle format (tf_bundle.py); re-parameterised variables and optimizer slots
    as a TF1 Saver would store them are handled on load."""
    from sga_b200 import tf_bundle, weights as Wm
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283                      # CRC-32C check value
    C = 8
    W = Wm.make_synthetic(C, 0)
    rn = Wm.runname(C, 0.01)
    d = tmp_path / rn
    d.mkdir()
    stored = {}
    for k, v in W.items():
        leaf = k.rsplit("/", 1)[-1]
        if leaf == "beta":
            stored[k.replace("/beta", "/reparam_beta")] = Wm.gdn_to_reparam(v, Wm.GDN_BETA_MIN)
        elif leaf == "gamma":
            stored[k.replace("/gamma", "/reparam_gamma")] = Wm.gdn_to_reparam(v, Wm.GDN_GAMMA_MIN)
        elif leaf == "kernel" and not k.startswith("mb_t2018hyper_synthesis_transform"):
            kh, kw, ci, co = v.shape                                          # default RDFT parameterisation (nn_models.py)
            M = Wm.irdft_matrix((kh, kw))
            stored[k + "_rdft"] = (M.T @ v.reshape(kh * kw, ci * co).astype(np.float64)).astype(np.float32)
        else:
            stored[k] = v
    stored["global_step"] = np.asarray(1000000, dtype=np.int64)
    stored["analysis_transform/layer_0/bias/Adam"] = np.zeros(C, np.float32)  # optimizer slot: ignored
    prefix = str(d / "model.ckpt-1000000")
    tf_bundle.write_bundle(prefix, stored, restart_interval=4, block_entries=7)   # several blocks, shared key prefixes
    assert tf_bundle.latest_checkpoint(str(d)) == prefix
    back = tf_bundle.read_bundle(prefix, verify=True)
    assert set(back) == set(stored)
    assert all(np.array_equal(back[k], stored[k]) and back[k].dtype == stored[k].dtype for k in stored)
    W2 = Wm.load_checkpoint(str(tmp_path), rn, C)
    assert set(W2) == set(W)
    for k in W:
        assert np.allclose(W2[k], W[k], rtol=1e-5, atol=1e-6), k
    bad = bytearray(open(prefix + ".index", "rb").read())
    bad[10] ^= 0xFF
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        tf_bundle.read_index(prefix + ".index")


def test_msssim_value_against_straight_line_numpy():
    """sga.py:175-176: tf.image.ssim_multiscale(x_hat, 255 x, 255).  msssim.py (torch, grouped convolutions) against an
    independent straight-line numpy implementation in the variance / covariance form, on a 192x208 pair (five scales, one
    odd-sized level that needs the symmetric pad) and on an identical pair (MS-SSIM = 1)."""
    import torch
    from sga_b200.msssim import msssim_per_image

    def gauss_window(a, w1d):
        # separable valid-mode filtering with explicit shifts
        h = sum(w1d[k] * a[k:a.shape[0] - len(w1d) + 1 + k, :] for k in range(len(w1d)))
        return sum(w1d[k] * h[:, k:h.shape[1] - len(w1d) + 1 + k] for k in range(len(w1d)))

    def ref_msssim(xh, xr):                                   # [H,W,3] float64, 0..255
        t = np.arange(11) - 5.0
        w1d = np.exp(-t * t / (2 * 1.5 ** 2)); w1d /= w1d.sum()
        c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
        pf = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]
        per_ch = []
        for ch in range(3):
            a, b = xh[..., ch], xr[..., ch]
            vals = []
            for lvl in range(5):
                if lvl > 0:
                    if a.shape[0] % 2: a, b = np.vstack([a, a[-1:]]), np.vstack([b, b[-1:]])
                    if a.shape[1] % 2: a, b = np.hstack([a, a[:, -1:]]), np.hstack([b, b[:, -1:]])
                    a = 0.25 * (a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])
                    b = 0.25 * (b[0::2, 0::2] + b[1::2, 0::2] + b[0::2, 1::2] + b[1::2, 1::2])
                mu_a, mu_b = gauss_window(a, w1d), gauss_window(b, w1d)
                var_a = gauss_window(a * a, w1d) - mu_a ** 2
                var_b = gauss_window(b * b, w1d) - mu_b ** 2
                cov = gauss_window(a * b, w1d) - mu_a * mu_b
                cs = (2 * cov + c2) / (var_a + var_b + c2)
                lum = (2 * mu_a * mu_b + c1) / (mu_a ** 2 + mu_b ** 2 + c1)
                vals.append(max((lum * cs).mean(), 0.0) if lvl == 4 else max(cs.mean(), 0.0))
            per_ch.append(np.prod([v ** p for v, p in zip(vals, pf)]))
        return float(np.mean(per_ch))

    rng = np.random.default_rng(3)
    h, w = 200, 360                                           # level sizes 200,100,50,25(odd -> padded),13
    base = rng.random((h // 8, w // 8, 3))
    x = np.kron(base, np.ones((8, 8, 1)))[:h, :w] * 0.8 + 0.1 * rng.random((h, w, 3))
    x = np.round(x * 255) / 255.0
    xt = x + 0.06 * rng.standard_normal(x.shape)
    got = msssim_per_image(torch.from_numpy(xt[None].astype(np.float32)), torch.from_numpy(x[None].astype(np.float32)))
    xh = np.round(np.clip(xt.astype(np.float32), 0, 1) * np.float32(255)).astype(np.float64)
    want = ref_msssim(xh, x.astype(np.float32).astype(np.float64) * 255.0)
    assert 0.3 < want < 0.999
    assert abs(float(got[0]) - want) < 2e-6, (float(got[0]), want)
    same = msssim_per_image(torch.from_numpy(x[None].astype(np.float32)), torch.from_numpy(x[None].astype(np.float32)))
    assert abs(float(same[0]) - 1.0) < 1e-6


def test_rank_segments_cover_every_image_once_and_stay_inside_one_reference_batch():
    """Cross-batch sharding (Tecnick, BASELINE configs[3]): 100 images of 1200x1200 are 15 reference batches (7,...,7,2;
    configs.py:5-9); every rank's contiguous global range is cut at the batch boundaries."""
    from sga_b200.dist import rank_segments, reference_batches, shard_range
    for n, px, ws in ((100, 1200 * 1200, 8), (100, 1200 * 1200, 3), (24, 768 * 512, 8), (5, 1200 * 1200, 8), (30, 768 * 512, 4)):
        batches = reference_batches(n, px)
        seen = np.zeros(n, dtype=int)
        for r in range(ws):
            segs = rank_segments(n, px, r, ws)
            lo, hi = shard_range(n, r, ws)
            assert sum(s["hi"] - s["lo"] for s in segs) == hi - lo
            for s in segs:
                b_lo, b_hi = batches[s["batch"]]
                assert b_lo <= s["lo"] < s["hi"] <= b_hi and s["n_ref"] == b_hi - b_lo and s["first_in_batch"] == s["lo"] - b_lo
                seen[s["lo"]:s["hi"]] += 1
        assert (seen == 1).all()
    per_gpu = [sum(s["hi"] - s["lo"] for s in rank_segments(100, 1200 * 1200, r, 8)) for r in range(8)]
    assert max(per_gpu) == 13 and min(per_gpu) == 12          # every GPU busy with its whole share at once
