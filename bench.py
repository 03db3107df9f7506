"""bench.py -- SGA iterations/sec on the Kodak-shaped synthetic batch (BASELINE.json: 24x768x512x3, 192 filters,
lambda 0.01), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one SGA iteration of the whole 24-image batch: Gumbel soft round of y,z -> h_s / g_s forward -> rate +
distortion -> backward to the latents -> Adam (reference sga.py:211-215).  The batch is sharded by image across the
ranks (no collective on the data path); the job is fixed as N grows => "scaling": "strong".

Rank 0 prints ONE JSON line.  `value` is device-resident throughput; `e2e` re-measures the same loop through the
public engine API with the image batch coming from pinned HOST memory every step (sga.py:212-214 feeds x every
iteration) and the step's statistics read back to the host; `roofline` is the tcgen05 (de)conv kernels' achieved
TFLOP/s from CUDA events around each of their launches in a second, instrumented pass of the same K steps; `cpu_baseline` is the CPU oracle (a PyTorch-CPU restatement of
the reference graph -- TF 1.15 is not installable) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_IMAGES, HEIGHT, WIDTH, C, LMBDA = 24, 768, 512, 192, 0.01
METRIC = "SGA iters/sec (768x512, 192 filters)"
UNIT = "it/s"


# ---------------------------------------------------------------------------------------------------
# algorithmic work: 2*MAC, transposed conv counted as in_px*Cin*Cout*25, fwd + dgrad, no wgrad
# ---------------------------------------------------------------------------------------------------
def flops_per_iteration(n, h, w, c):
    hy, wy = -(-h // 16), -(-w // 16)
    hz, wz = -(-(-(-hy // 2)) // 2), -(-(-(-wy // 2)) // 2)
    mid = int(c * 1.5)
    f = {}
    px = n * hz * wz
    f["hs0"] = 2.0 * px * c * c * 25
    f["hs1"] = 2.0 * (4 * px) * c * mid * 25
    f["hs2"] = 2.0 * (16 * px) * mid * (2 * c) * 9
    py = n * hy * wy
    for i in range(3):
        f["gs%d" % i] = 2.0 * (py * 4 ** i) * c * c * 25
        f["igdn%d" % i] = 2.0 * (py * 4 ** (i + 1)) * c * c
    f["gs3"] = 2.0 * (py * 64) * c * 3 * 25
    fwd = sum(f.values())
    return f, 2.0 * fwd


def measured_traffic():
    """DRAM bytes of the (de)conv launches of one step from the newest committed `ncu --set full` capture
    (profiles/*_traffic.json, written from the .ncu-rep by tools/ncu_traffic.py)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    with open(files[-1]) as fh:
        d = json.load(fh)
    return d.get("conv_dram_bytes_per_launch"), os.path.basename(files[-1])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    hbm=float(d.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.proc = None
        self.path = "/tmp/sga_bench_clocks_%d.csv" % os.getpid()
        sel = str(device_index)
        try:
            import torch
            u = str(torch.cuda.get_device_properties(device_index).uuid)
            sel = u if u.startswith("GPU-") else "GPU-" + u
        except Exception:
            pass
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", sel, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------------
# the CPU arm: oracle port of the reference graph, driven like sga.py:210-215
# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, budget_s=150.0, samples=((768, 512), (384, 256), (256, 256), (128, 128))):
    """Returns (it/s extrapolated to the 24x768x512 batch, cores, sample description, seconds per sample step)."""
    import torch
    import sga_b200
    from oracle import sga_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W = O._to(sga_b200.weights.make_synthetic(C, 0), torch.float32)

    def run(h, w, n_steps, n_warm):
        x = torch.from_numpy(sga_b200.synth.synthetic_images(1, h, w, 0).astype(np.float32) / 255.0)
        y, z = O.analysis(x, W)
        opt = O.Adam(lr=0.005)
        ts = []
        for it in range(n_warm + n_steps):
            t0 = time.perf_counter()
            T = O.annealed_temperature(it, r=1e-3, ub=0.5, scheme="exp0", t0=700)
            out = O.sga_graph(x, y, z, W, T, O.draw_noise(y, z, 0, it), LMBDA, n_ref=N_IMAGES)
            y, z = opt.update([y, z], [out["grad_y"], out["grad_z"]])
            _ = float(out["rd_loss"])
            if it >= n_warm:
                ts.append(time.perf_counter() - t0)
        return float(np.mean(ts))

    t_probe = run(128, 128, 1, 1)                       # also warms the CPU kernels
    per_px = t_probe / (128 * 128)
    chosen = samples[-1]
    for (h, w) in samples:
        if per_px * h * w * (steps + warmup) <= budget_s:
            chosen = (h, w)
            break
    t = run(chosen[0], chosen[1], steps, warmup)
    scale = (N_IMAGES * HEIGHT * WIDTH) / float(chosen[0] * chosen[1])
    sample = ("1 image of %dx%d (1/%.0f of the 24x768x512 batch by pixels), %d timed SGA iterations; PyTorch-CPU "
              "restatement of the reference TF1 graph (oracle/sga_oracle.py), it/s extrapolated linearly in pixels"
              % (chosen[0], chosen[1], scale, steps))
    return 1.0 / (t * scale), cores, sample, t


def main_reference(args, rank):
    if rank != 0:
        return
    v, cores, sample, t = cpu_reference_rate(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SGA 24x768x512x3, num_filters=192, lambda=0.01 (BASELINE configs[1])"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def main_ours(args, rank, ws, local):
    import torch
    import torch.distributed as dist
    import sga_b200
    from sga_b200 import dist as D
    from sga_b200.engine import SgaEngine
    from sga_b200.schedule import annealed_temperature

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lo, hi = D.shard_range(N_IMAGES, rank, ws)
    n_loc = hi - lo
    W = sga_b200.weights.make_synthetic(C, 0)
    imgs = sga_b200.synth.synthetic_images(N_IMAGES, HEIGHT, WIDTH, 0)[lo:hi]
    x_host = torch.from_numpy(imgs.astype(np.float32) / 255.0).pin_memory()
    x = x_host.to(dev)
    eng = SgaEngine(W, n_loc, HEIGHT, WIDTH, C, LMBDA, n_ref=N_IMAGES, operand=args.operand, seed=0)
    y0, z0 = eng.encode_init(x)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if ws == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    temp = lambda it: annealed_temperature(it, r=1e-3, ub=0.5, scheme="exp0", t0=700)

    # ---- (1) device-resident loop: `value` (un-instrumented), then the same K steps again with one CUDA event pair
    #      around every launch on the launching stream for the roofline (the event pairs serialise the launches, i.e.
    #      switch off the programmatic-dependent-launch overlap, so they stay out of the pass that yields `value`)
    st = eng.new_state(y0.clone(), z0.clone())
    for it in range(args.warmup):
        eng.step(x, st, temp(it), it, want_stats=False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        it = args.warmup + k
        eng.step(x, st, temp(it), it, want_stats=False)
    e1.record()
    barrier()
    launches = eng.launch_count() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = 1000.0 / ms_per_step
    eng.profile(True)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for k in range(args.steps):
        it = args.warmup + args.steps + k
        eng.step(x, st, temp(it), it, want_stats=False)
    e3.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    prof = eng.profile_read()
    eng.profile(False)
    ms_per_step_instrumented = max_over_ranks(e2.elapsed_time(e3)) / args.steps

    # ---- (2) end-to-end through the public API with host buffers
    st2 = eng.new_state(y0.clone(), z0.clone())
    xbuf = [torch.empty_like(x), torch.empty_like(x)]
    copy_stream = torch.cuda.Stream()
    stats_host = torch.empty((args.steps + args.warmup, eng.n_stats), dtype=torch.float32).pin_memory()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_steps(first, count):
        main = torch.cuda.current_stream()
        for k in range(count):
            it = first + k
            b = k & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[b])
                xbuf[b].copy_(x_host, non_blocking=True)          # host -> device every step (sga.py:212-214)
                ready[b].record(copy_stream)
            main.wait_event(ready[b])
            eng.step(xbuf[b], st2, temp(it), it, want_stats=True)
            freed[b].record(main)
            stats_host[it].copy_(eng.stats, non_blocking=True)    # device -> host read of the step's statistics
        torch.cuda.current_stream().synchronize()

    for b in range(2):
        freed[b].record(torch.cuda.current_stream())
    e2e_steps(0, args.warmup)
    barrier()
    e0.record()
    e2e_steps(args.warmup, args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    last = eng.unpack_stats(stats_host[args.warmup + args.steps - 1].numpy())
    if not (np.isfinite(stats_host.numpy()).all()):
        raise RuntimeError("non-finite SGA statistics inside the timed loop (rd_loss %r): the measurement is void" % last["rd_loss"])

    if rank != 0:
        return
    fl, flops_iter = flops_per_iteration(n_loc, HEIGHT, WIDTH, C)
    pk = peaks()
    conv_labels = {"hs0_fwd": "hs0", "hs1_fwd": "hs1", "hs2_fwd": "hs2", "hs0_dgrad": "hs0", "hs1_dgrad": "hs1",
                   "hs2_dgrad": "hs2", "gs0_fwd": "gs0", "gs1_fwd": "gs1", "gs2_fwd": "gs2", "gs0_dgrad": "gs0",
                   "gs1_dgrad": "gs1", "gs2_dgrad": "gs2"}
    # algorithmic FLOPs executed inside each (de)conv launch: the conv itself plus the IGDN channel mixing fused into it
    # (forward: IGDN_i after conv_i; backward: IGDN_{i-1} in the dgrad of conv_i; IGDN_2's backward is in k_l3_dgrad_igdn)
    fused_extra = {"gs0_fwd": "igdn0", "gs1_fwd": "igdn1", "gs2_fwd": "igdn2", "gs1_dgrad": "igdn0", "gs2_dgrad": "igdn1"}
    kfl = {k: fl[conv_labels[k]] + (fl[fused_extra[k]] if (k in fused_extra and C in (64, 128, 192)) else 0.0)
           for k in conv_labels}
    conv_ms = sum(prof[k][0] for k in conv_labels if k in prof)
    conv_fl = sum(kfl[k] * prof[k][1] for k in conv_labels if k in prof)
    conv_launches = sum(prof[k][1] for k in conv_labels if k in prof)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_src = measured_traffic()
    if traffic is not None and ws > 1:
        traffic = traffic / ws          # the capture is of the 24-image job on one GPU
    breakdown = {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items())}
    per_layer = {k: round(kfl[k] * prof[k][1] / (prof[k][0] * 1e-3) / 1e12, 1) for k in conv_labels
                 if k in prof and prof[k][0] > 0}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.operand == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "SGA 24x768x512x3, num_filters=192, lambda=0.01 (BASELINE configs[1])",
                   "images_per_gpu": n_loc, "operand": args.operand + " operands, fp32 accumulate (TMEM); latents/Adam fp32",
                   "cache": "inputs larger than L2 (activations %.1f GB per step)" % (eng.workspace_bytes / 1e9),
                   "final_step": {"rd_loss": last["rd_loss"], "train_bpp": last["train_bpp"], "train_mse": last["train_mse"]}},
        "e2e": {"value": 1000.0 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4 * ws),
                "d2h_bytes_per_step": int(eng.n_stats * 4 * ws), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor",
                     "kernel": "tcgen05 implicit-GEMM (de)conv kernels: k_conv_tc x9 + k_conv_igdn_fwd x3 per step",
                     "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                     "traffic": traffic, "traffic_unit": "DRAM bytes per launch (mean over the 12 launches, 24 images)",
                     "traffic_source": traffic_src, "peak_source": pk["source"], "launches_timed": conv_launches,
                     "timed": "second pass of the same K steps with a CUDA event pair around every launch (%.3f ms/step; "
                              "the value pass is un-instrumented)" % ms_per_step_instrumented,
                     "per_layer_tflops": per_layer,
                     "whole_step_tflops": flops_iter * value / 1e12,
                     "whole_step_frac": flops_iter * value / 1e12 / pk["tflops"]},
        "ms_per_step_by_kernel": breakdown,
        "clocks": clocks,
    }
    if ws == 1 and not args.no_cpu:
        v, cores, sample, _ = cpu_reference_rate(2, 1, budget_s=40.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--operand", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank)
        return
    import __graft_entry__ as ge
    if rank == 0 or not os.path.exists(os.path.join(ge.PKG, "libsga_b200.so")):
        ge.build()
    if ws > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    main_ours(args, rank, ws, local)
    if ws > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
