"""BASELINE configs[3]: Tecnick-shaped synthetic input (100 x 1200 x 1200, reference batches of 7, configs.py:5-9) over the
GPUs of one box, sweeping the temperature-annealing schedule (--annealing_rate x --t0, tf_boilerplate.py:171-176,
utils.py:151-180).  One process per GPU (torchrun); every rank holds the engines of its segments (dist.rank_segments) for
the whole sweep, so the 16 schedules cost one start-up.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/tecnick_sweep.py \
        --its 500 --out gpurun_out/tecnick_sweep.json
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
    ap.add_argument("--images", type=int, default=100)
    ap.add_argument("--size", type=int, default=1200)
    ap.add_argument("--its", type=int, default=500)
    ap.add_argument("--rates", default="5e-4,1e-3,2e-3,4e-3")
    ap.add_argument("--t0s", default="0,350,700,1000")
    ap.add_argument("--out", default="gpurun_out/tecnick_sweep.json")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    import sga_b200
    from sga_b200 import dist as D
    from sga_b200.engine import SgaEngine
    from sga_b200.schedule import annealed_temperature
    C, H, W, lmbda = 192, a.size, a.size, 0.01
    Wt = sga_b200.weights.make_synthetic(C, 0)
    segs = D.rank_segments(a.images, H * W, rank, ws)
    jobs = []
    for sg in segs:
        n = sg["hi"] - sg["lo"]
        x = torch.from_numpy(sga_b200.synth.synthetic_images(n, H, W, sg["lo"]).astype(np.float32) / 255.0).cuda()
        eng = SgaEngine(Wt, n, H, W, C, lmbda, n_ref=sg["n_ref"], operand="bf16", seed=sg["batch"], first_image=sg["first_in_batch"])
        jobs.append(dict(eng=eng, x=x, init=eng.encode_init(x)))
    results = []
    for r in [float(v) for v in a.rates.split(",")]:
        for t0 in [int(v) for v in a.t0s.split(",")]:
            temps = [annealed_temperature(it, r=r, ub=0.5, scheme="exp0", t0=t0) for it in range(a.its)]
            states = [j["eng"].new_state(j["init"][0].clone(), j["init"][1].clone()) for j in jobs]
            torch.cuda.synchronize()
            if ws > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s0 in range(0, a.its, 8):                      # all reference batches of a rank advance together
                for j, st in zip(jobs, states):
                    j["eng"].loop(j["x"], st, temps[s0:s0 + 8], s0, want_stats=False)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            rows = []
            for j, st in zip(jobs, states):
                ev = j["eng"].eval_hard(j["x"], st["y"], st["z"])
                rows.append(np.stack([ev["est_bpp"], ev["psnr"]], axis=1))
            rows = np.concatenate(rows) if rows else np.zeros((0, 2))
            full = D.gather_rows(rows, a.images, rank, ws)
            if ws > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                rec = dict(annealing_rate=r, t0=t0, T_final=temps[-1], its=a.its, seconds=float(ms.item()) / 1e3,
                           its_per_s=a.its / (float(ms.item()) / 1e3), avg_est_bpp=float(full[:, 0].mean()),
                           avg_psnr=float(full[:, 1].mean()))
                results.append(rec)
                print("r=%g t0=%d: %.1f it/s (%.2f s for %d its of %d images), avg bpp %.4f psnr %.3f" % (
                    r, t0, rec["its_per_s"], rec["seconds"], a.its, a.images, rec["avg_est_bpp"], rec["avg_psnr"]), flush=True)
    if rank == 0:
        out = dict(workload="SGA Tecnick-shaped %dx%dx%dx3, num_filters=192, lambda=0.01, %d iterations per schedule (the reference "
                            "runs 2000; throughput does not depend on the schedule)" % (a.images, H, W, a.its),
                   n_gpus=ws, images_per_gpu=[sum(s["hi"] - s["lo"] for s in D.rank_segments(a.images, H * W, r_, ws)) for r_ in range(ws)],
                   reference_batches=len(D.reference_batches(a.images, H * W)), sweep=results)
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
