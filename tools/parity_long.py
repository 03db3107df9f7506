"""Long-run parity of the benchmarked bf16 loop against the fp32 loop (north_star tolerance: 1e-3 bpp, 0.01 dB).

Runs the whole compress() body of the reference (sga.py:200-247: init, `its` SGA iterations, hard-rounded evaluation in
fp32) on the same images / weights / Philox noise seed once per arm and writes the per-image final est_bpp / PSNR of
every arm plus the deltas against the first arm to a JSON file.

    python tools/parity_long.py --arms f32:0,bf16:0,f32:1 --n 2 --its 2000 --out gpurun_out/parity.json

An arm is <operand>:<seed>[:ENV=VAL[+ENV=VAL...]]; the environment assignments are applied while that arm's engine is
created and run (the library reads its tuning switches with getenv at plan / launch time).  The pseudo-variable
PERTURB=<rel> is not exported: it multiplies the initial y by (1 + rel * N(0,1)) -- the control arm that separates the
chaotic divergence of a 2000-step stochastic optimisation from the effect of the operand precision.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arms", default="f32:0,bf16:0,f32:1")
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--h", type=int, default=768)
    ap.add_argument("--w", type=int, default=512)
    ap.add_argument("--c", type=int, default=192)
    ap.add_argument("--lmbda", type=float, default=0.01)
    ap.add_argument("--its", type=int, default=2000)
    ap.add_argument("--annealing_rate", type=float, default=1e-3)
    ap.add_argument("--t0", type=int, default=700)
    ap.add_argument("--weights_seed", type=int, default=0)
    ap.add_argument("--eval_at", default="50,100,200,500,1000", help="extra hard-rounded evaluations after these step counts")
    ap.add_argument("--out", default="gpurun_out/parity_long.json")
    a = ap.parse_args()

    import torch
    import __graft_entry__ as ge
    ge.build()
    import sga_b200
    from sga_b200.engine import SgaEngine
    from sga_b200.schedule import annealed_temperature
    eval_at = set(int(v) for v in a.eval_at.split(",") if v)

    W = sga_b200.weights.make_synthetic(a.c, a.weights_seed)
    x = torch.from_numpy(sga_b200.synth.synthetic_images(a.n, a.h, a.w, 0).astype(np.float32) / 255.0).cuda()
    arms = []
    for spec in a.arms.split(","):
        parts = spec.split(":")
        env = dict(kv.split("=") for kv in parts[2].split("+")) if len(parts) > 2 and parts[2] else {}
        arms.append(dict(name=spec, operand=parts[0], seed=int(parts[1]), env=env))
    out = dict(config=dict(n=a.n, h=a.h, w=a.w, c=a.c, lmbda=a.lmbda, its=a.its, annealing_rate=a.annealing_rate, t0=a.t0,
                           weights="synthetic seed %d" % a.weights_seed, images="synthetic seed 0"), arms=[])
    for arm in arms:
        perturb = float(arm["env"].pop("PERTURB", 0.0))
        old = {k: os.environ.get(k) for k in arm["env"]}
        os.environ.update(arm["env"])
        eng = SgaEngine(W, a.n, a.h, a.w, a.c, a.lmbda, operand=arm["operand"], seed=arm["seed"])
        traj, checkpoints = [], {}
        torch.cuda.synchronize()
        t0 = time.time()
        y, z = eng.encode_init(x)
        if perturb > 0.0:       # control arm: the same arithmetic started from a relatively perturbed y (chaos, not precision)
            g = torch.Generator(device="cuda").manual_seed(1234)
            y = y * (1.0 + perturb * torch.randn(y.shape, generator=g, device="cuda"))
        st = eng.new_state(y, z)
        for it in range(a.its):
            T = annealed_temperature(it, r=a.annealing_rate, ub=0.5, scheme="exp0", t0=a.t0)
            log_pt = it % 100 == 0 or it + 1 == a.its
            eng.step(x, st, T, it, want_stats=log_pt)
            if log_pt:
                s_ = eng.unpack_stats(eng.stats.cpu().numpy())
                traj.append([it, s_["rd_loss"], s_["train_bpp"], s_["train_mse"]])
            if (it + 1) in eval_at:
                e_ = eng.eval_hard(x, st["y"], st["z"])
                checkpoints[it + 1] = dict(est_bpp=e_["est_bpp"].tolist(), psnr=e_["psnr"].tolist())
        y, z = st["y"], st["z"]
        ev = eng.eval_hard(x, y, z)
        torch.cuda.synchronize()
        dt = time.time() - t0
        eng.close()
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        rec = dict(name=arm["name"], seconds=dt, est_bpp=ev["est_bpp"].tolist(), psnr=ev["psnr"].tolist(),
                   mse=ev["mse"].tolist(), est_y_bpp=ev["est_y_bpp"].tolist(), est_z_bpp=ev["est_z_bpp"].tolist(),
                   y_absmax=float(y.abs().max()), trajectory=traj, checkpoints=checkpoints)
        out["arms"].append(rec)
        print("%-28s %.1fs  bpp %s  psnr %s" % (arm["name"], dt, np.round(ev["est_bpp"], 5), np.round(ev["psnr"], 4)), flush=True)
    base = out["arms"][0]
    for rec in out["arms"][1:]:
        rec["d_bpp_vs_first"] = (np.asarray(rec["est_bpp"]) - np.asarray(base["est_bpp"])).tolist()
        rec["d_psnr_vs_first"] = (np.asarray(rec["psnr"]) - np.asarray(base["psnr"])).tolist()
        rec["d_checkpoints_vs_first"] = {
            k: dict(max_abs_dbpp=float(np.abs(np.asarray(v["est_bpp"]) - np.asarray(base["checkpoints"][k]["est_bpp"])).max()),
                    mean_dbpp=float((np.asarray(v["est_bpp"]) - np.asarray(base["checkpoints"][k]["est_bpp"])).mean()),
                    max_abs_dpsnr=float(np.abs(np.asarray(v["psnr"]) - np.asarray(base["checkpoints"][k]["psnr"])).max()))
            for k, v in rec["checkpoints"].items() if k in base["checkpoints"]}
        print("    growth: " + "  ".join("%d: %.1e/%.1e" % (k, d["max_abs_dbpp"], d["max_abs_dpsnr"])
                                         for k, d in sorted(rec["d_checkpoints_vs_first"].items())), flush=True)
        print("%-28s vs %s: max|dbpp| %.2e  max|dpsnr| %.2e dB" % (
            rec["name"], base["name"], np.abs(rec["d_bpp_vs_first"]).max(), np.abs(rec["d_psnr_vs_first"]).max()), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
