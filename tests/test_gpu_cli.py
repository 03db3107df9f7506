"""The `compress` command end to end on the GPU: reference-compatible flags, checkpoint layout, output file."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(tmp_path, C=64, n=2, h=64, w=96):
    import sga_b200
    rn = sga_b200.weights.runname(C, 0.01)
    sga_b200.weights.save_checkpoint(str(tmp_path / "ckpt"), rn, sga_b200.weights.make_synthetic(C, 0))
    inp = tmp_path / "imgs.npy"
    np.save(inp, sga_b200.synth.synthetic_images(n, h, w, 1))
    (tmp_path / "res").mkdir()
    return rn, inp


@pytest.mark.parametrize("operand", ["bf16", "f32"])
def test_compress_cli(tmp_path, operand):
    rn, inp = _setup(tmp_path)
    cmd = [sys.executable, os.path.join(ROOT, "sga.py"), "--num_filters", "64", "--checkpoint_dir", str(tmp_path / "ckpt"),
           "compress", "--sga_its", "12", "--results_dir", str(tmp_path / "res"), "--operand", operand, rn, str(inp)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "Defaulting lmbda (mse coefficient) to 0.01 as used in model training." in out      # sga.py:157-159
    assert "----RD Optimization----" in out and "it=0, T=0.500 rd_loss=" in out and "it=11, T=0.500" in out
    assert "Avg psnr:" in out and "Avg est_bpp:" in out
    f = tmp_path / "res" / ("rd-sga-lmbda=0.01+%s-input=imgs.npy.npz" % rn)                  # sga.py:263-268
    assert f.exists(), os.listdir(tmp_path / "res")
    d = np.load(f)
    assert sorted(d.files) == sorted(["mse", "psnr", "msssim", "msssim_db", "est_bpp", "est_y_bpp", "est_z_bpp"])
    assert all(d[k].shape == (2,) for k in d.files)
    assert np.allclose(d["est_bpp"], d["est_y_bpp"] + d["est_z_bpp"], rtol=1e-5)
    assert np.isfinite(d["psnr"]).all() and (d["est_bpp"] > 0).all()


def test_compress_cli_two_ranks_match_one(tmp_path):
    """Sharding the batch over 2 GPUs gives the same per-image results as one GPU (1/n_ref scale, same noise)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rn, inp = _setup(tmp_path, n=4)
    base = ["--num_filters", "64", "--checkpoint_dir", str(tmp_path / "ckpt"), "compress", "--sga_its", "8",
            "--operand", "f32"]
    res = []
    for tag, launcher in (("one", [sys.executable]),
                          ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29741"])):
        rd = tmp_path / ("res_" + tag)
        rd.mkdir()
        r = subprocess.run(launcher + [os.path.join(ROOT, "sga.py")] + base + ["--results_dir", str(rd), rn, str(inp)],
                           capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
        assert r.returncode == 0, r.stdout + r.stderr
        res.append(np.load(rd / ("rd-sga-lmbda=0.01+%s-input=imgs.npy.npz" % rn)))
    # the Philox counter is the element index inside the whole reference batch (sga_config.first_image) and the
    # gradient scale is 1/n_ref on every shard: the per-image results do not depend on the sharding
    assert np.abs(res[0]["est_bpp"] - res[1]["est_bpp"]).max() < 1e-4 * res[0]["est_bpp"].max()
    assert np.abs(res[0]["psnr"] - res[1]["psnr"]).max() < 1e-3


def test_compress_cli_bits_back(tmp_path):
    """bb_sga.py compress: both loops, the 8 result fields of bb_sga.py:176, reference file naming."""
    import sga_b200
    rn = sga_b200.weights.runname(64, 0.01, prefix="mbt2018_bb")
    sga_b200.weights.save_checkpoint(str(tmp_path / "ckpt"), rn, sga_b200.weights.make_synthetic_bb(64, 0))
    inp = tmp_path / "imgs.npy"
    np.save(inp, sga_b200.synth.synthetic_images(2, 64, 96, 1))
    (tmp_path / "res").mkdir()
    cmd = [sys.executable, os.path.join(ROOT, "bb_sga.py"), "--num_filters", "64", "--checkpoint_dir", str(tmp_path / "ckpt"),
           "compress", "--sga_its", "12", "--rate_its", "8", "--results_dir", str(tmp_path / "res"), rn, str(inp)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "----RD Optimization----" in out and "it=0, T=0.500 rd_loss=" in out and "it=11, T=0.500" in out
    assert "----Rate Optimization----" in out and "it= 7" in out and "rate=" in out
    assert "Avg est_bpp_back:" in out
    f = tmp_path / "res" / ("rd-bb_sga-lmbda=0.01+%s-input=imgs.npy.npz" % rn)                   # bb_sga.py:283-287
    assert f.exists(), os.listdir(tmp_path / "res")
    d = np.load(f)
    assert sorted(d.files) == sorted(["mse", "psnr", "msssim", "msssim_db", "est_bpp", "est_y_bpp", "est_z_bpp",
                                      "est_bpp_back"])
    assert all(d[k].shape == (2,) for k in d.files)
    assert np.allclose(d["est_bpp"], d["est_y_bpp"] + d["est_z_bpp"] - d["est_bpp_back"], rtol=1e-4, atol=1e-5)
    assert np.isfinite(d["psnr"]).all()


def test_compress_cli_many_reference_batches_opt_record_and_reconstruction(tmp_path):
    """An input that spans several reference batches (configs.py:5-9; forced here with a small pixel budget): one engine per
    batch, all advancing together, the reference's per-batch log blocks in batch order, per-image results in input order and
    equal to what each batch gives when compressed on its own (n_ref and the noise key are per batch).  Also the optional
    outputs: opt-*.npz (sga.py:271-278) and, for a single image, recon-*.png (sga.py:280-291)."""
    rn, inp = _setup(tmp_path, n=5, h=64, w=64)
    env = dict(os.environ, SGA_EVAL_BATCH_NUM_PIXELS=str(2 * 64 * 64))      # batches of 2, 2, 1 images
    base = [sys.executable, os.path.join(ROOT, "sga.py"), "--num_filters", "64", "--checkpoint_dir", str(tmp_path / "ckpt"),
            "compress", "--sga_its", "25", "--operand", "f32", "--save_opt_record"]
    r = subprocess.run(base + ["--results_dir", str(tmp_path / "res"), rn, str(inp)], capture_output=True, text=True,
                       timeout=600, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("----RD Optimization----") == 3 and "SGA loop: 25 iterations of 5 images (3 reference batches)" in r.stdout
    assert r.stdout.count("it=10, T=0.500") == 3                       # log_itv = 10 with the optimisation record on
    full = np.load(tmp_path / "res" / ("rd-sga-lmbda=0.01+%s-input=imgs.npy.npz" % rn))
    rec = np.load(tmp_path / "res" / ("opt-sga-lmbda=0.01+%s-input=imgs.npy.npz" % rn))
    assert list(rec["its"]) == [0, 10, 20, 24] and rec["rd_loss"].shape == (4,) and np.isfinite(rec["rd_loss"]).all()
    # the last batch (image 4 alone, n_ref = 1, noise key 2) compressed on its own must differ only by the noise key:
    # compress images 0-1 alone (batch 0: same key 0, same n_ref 2) and compare exactly
    X = np.load(inp)
    np.save(tmp_path / "first.npy", X[:2])
    (tmp_path / "res1").mkdir()
    r1 = subprocess.run(base + ["--results_dir", str(tmp_path / "res1"), rn, str(tmp_path / "first.npy")], capture_output=True,
                        text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert r1.returncode == 0, r1.stdout + r1.stderr
    first = np.load(tmp_path / "res1" / ("rd-sga-lmbda=0.01+%s-input=first.npy.npz" % rn))
    for k in ("psnr", "est_bpp", "mse"):
        assert np.allclose(full[k][:2], first[k], rtol=1e-5, atol=1e-6), k
    # single image: reconstruction PNG
    from PIL import Image
    Image.fromarray(X[0]).save(tmp_path / "one.png")
    (tmp_path / "res2").mkdir()
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "sga.py"), "--num_filters", "64", "--checkpoint_dir", str(tmp_path / "ckpt"),
                         "compress", "--sga_its", "6", "--save_reconstruction", "--results_dir", str(tmp_path / "res2"), rn,
                         str(tmp_path / "one.png")], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r2.returncode == 0, r2.stdout + r2.stderr
    png = tmp_path / "res2" / ("recon-sga-lmbda=0.01-rd_opt_its=6+%s-input=one.png.png" % rn)
    assert png.exists() and np.asarray(Image.open(png)).shape == (64, 64, 3)


def test_compress_cli_two_ranks_many_batches_match_one(tmp_path):
    """Cross-batch sharding over 2 GPUs: 5 images in reference batches of 2, 2, 1 (forced by a small pixel budget), rank 0
    holds images 0-2 (batch 0 and half of batch 1), rank 1 images 3-4 (the other half and batch 2).  Per-image results equal
    the one-GPU run: n_ref, noise key and noise offset are those of the image's reference batch, not of the shard."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rn, inp = _setup(tmp_path, n=5, h=64, w=64)
    env = dict(os.environ, SGA_EVAL_BATCH_NUM_PIXELS=str(2 * 64 * 64))
    base = ["--num_filters", "64", "--checkpoint_dir", str(tmp_path / "ckpt"), "compress", "--sga_its", "12", "--operand", "f32"]
    res = []
    for tag, launcher in (("one", [sys.executable]),
                          ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29743"])):
        rd = tmp_path / ("res_" + tag)
        rd.mkdir()
        r = subprocess.run(launcher + [os.path.join(ROOT, "sga.py")] + base + ["--results_dir", str(rd), rn, str(inp)],
                           capture_output=True, text=True, timeout=900, cwd=str(tmp_path), env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("----RD Optimization----") == 3
        res.append((np.load(rd / ("rd-sga-lmbda=0.01+%s-input=imgs.npy.npz" % rn)), r.stdout))
    for k in ("psnr", "est_bpp", "est_y_bpp", "est_z_bpp", "mse"):
        assert res[0][0][k].shape == (5,)
        assert np.allclose(res[0][0][k], res[1][0][k], rtol=1e-5, atol=1e-6), (k, res[0][0][k], res[1][0][k])
    # the per-batch log lines (batch-wide sums re-assembled over the ranks) agree too
    import re
    lines = [[l for l in out.splitlines() if l.startswith("it=11,")] for _, out in res]
    assert len(lines[0]) == 3 and len(lines[1]) == 3, lines
    nums = [np.array([[float(v) for v in re.findall(r"=(-?[0-9.]+)", l)] for l in ls]) for ls in lines]
    assert np.allclose(nums[0], nums[1], rtol=2e-4, atol=2e-4), lines


def test_compress_cli_writes_the_bitstream_of_the_sga_result(tmp_path):
    """With an output file, sga.py writes the range-coded round(y), round(z) (sga.py:240-241) and reports the actual rate."""
    from sga_b200 import codec as Cd
    rn, inp = _setup(tmp_path, n=2, h=64, w=96)
    out_file = tmp_path / "out.sgac"
    cmd = [sys.executable, os.path.join(ROOT, "sga.py"), "--num_filters", "64", "--checkpoint_dir", str(tmp_path / "ckpt"),
           "compress", "--sga_its", "6", "--results_dir", str(tmp_path / "res"), rn, str(inp), str(out_file)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Avg actual_bpp:" in r.stdout
    streams, mode, C, H, W = Cd.unpack(out_file.read_bytes())
    assert (mode, C, H, W) == (Cd.MODE_SGA, 64, 64, 96) and len(streams) == 2
    d = np.load(tmp_path / "res" / ("rd-sga-lmbda=0.01+%s-input=imgs.npy.npz" % rn))
    assert d["actual_bpp"].shape == (2,) and (d["actual_bpp"] > 0).all()
