"""bench.py -- SGA iterations/sec on the reference's workloads, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload kodak192|kodak256|tecnick|bb]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Default workload = BASELINE.json configs[1]: the Kodak-shaped synthetic batch 24x768x512x3, 192 filters, lambda 0.01.
A "step" is one SGA iteration of the whole batch: Gumbel soft round of y,z -> h_s / g_s forward -> rate + distortion ->
backward to the latents -> Adam (reference sga.py:211-215).  The batch is sharded by image across the ranks (no collective
on the data path); the job is fixed as N grows => "scaling": "strong".

Rank 0 prints ONE JSON line.
  value        device-resident throughput of K iterations through sga_loop (the call the CLI makes: the two-stream step
               issued K times from inside the library), CUDA events on the launching stream, max over ranks;
  e2e          the same loop through the public engine API one iteration per call, with the image batch copied from
               pinned HOST memory every iteration (sga.py:212-214 feeds x every iteration) and the iteration's statistics
               read back to the host;
  roofline     the DOMINANT kernel (g_s layer 2 forward: transposed conv + fused inverse GDN): algorithmic FLOPs per launch /
               its mean launch duration from CUDA event pairs in a third, instrumented pass of the same K iterations;
               `kernels` lists every launch group of the step the same way (tensor-bound ones against the measured bf16
               peak, the two layer-3 kernels against the measured HBM bandwidth), `whole_step` the whole iteration;
  cpu_baseline the CPU restatement of the reference loop (baseline/cpu_sga.py: autograd backward, host Adam, x re-fed
               every iteration) on the SAME batch, all host cores, 2 timed iterations (N = 1 only);
  --impl reference   times that CPU arm alone, K timed iterations of the full batch.
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

UNIT = "it/s"
WORKLOADS = {
    # name: images, height, width, filters, lambda, reference batch size, variant, metric label, description
    "kodak192": dict(n=24, h=768, w=512, c=192, lmbda=0.01, variant="sga", metric="SGA iters/sec (768x512, 192 filters)",
                     desc="SGA 24x768x512x3, num_filters=192, lambda=0.01 (BASELINE configs[1])"),
    "kodak256": dict(n=24, h=768, w=512, c=256, lmbda=0.08, variant="sga", metric="SGA iters/sec (768x512, 256 filters)",
                     desc="SGA 24x768x512x3, num_filters=256, lambda=0.08 (BASELINE configs[2], high-rate model)"),
    "tecnick": dict(n=100, h=1200, w=1200, c=192, lmbda=0.01, variant="sga", metric="SGA iters/sec (1200x1200, 192 filters)",
                    desc="SGA Tecnick-shaped 100x1200x1200x3 in reference batches of 7 (configs.py:5-9), num_filters=192 "
                         "(BASELINE configs[3])"),
    "bb": dict(n=24, h=768, w=512, c=192, lmbda=0.01, variant="bb", metric="bb_sga R-D iters/sec (768x512, 192 filters)",
               desc="bb_sga.py bits-back variant 24x768x512x3, num_filters=192, lambda=0.01 (BASELINE configs[4])"),
}
TRAFFIC_FILE = os.path.join("profiles", "r2_traffic.json")     # per-kernel DRAM bytes from the committed ncu --set full capture
TECNICK_PER_GPU = 13                                           # ceil(100 / 8): the 8-GPU share; fewer GPUs run 13 x N images


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


def layer3_bytes(n, h, w, c):
    """Algorithmic HBM bytes of the two layer-3 kernels: bf16 act_2 / s_2 / gb_2 [n,8hy,8wy,c], fp32 x,
    bf16 gx4 [n,16hy,16wy,4]."""
    hy, wy = -(-h // 16), -(-w // 16)
    act = n * (8 * hy) * (8 * wy) * c * 2
    gx4 = n * (16 * hy) * (16 * wy) * 4 * 2
    return {"l3_fwd": act + n * h * w * 3 * 4 + gx4, "l3_dgrad": gx4 + 2 * act + act}


def measured_traffic():
    p = os.path.join(ROOT, TRAFFIC_FILE)
    if not os.path.exists(p):
        return {}
    with open(p) as fh:
        return json.load(fh)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    hbm=float(d.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json: sustained bf16, copy bandwidth)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    # NVML reasons bit mask (nvml.h: nvmlClocksEventReason*)
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def _start_nvml(self, device_index):
        """In-process NVML polling thread (one sample every ~2 ms): a multi-GPU run's timed region is only tens of
        milliseconds long, less than it takes an `nvidia-smi -lms` child process to print its first line."""
        import threading
        import pynvml
        import torch
        pynvml.nvmlInit()
        u = str(torch.cuda.get_device_properties(device_index).uuid)
        u = u if u.startswith("GPU-") else "GPU-" + u
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(u.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID(u)
        self.nvml_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        self.nvml_sm, self.nvml_mask, self.nvml_stop = [], [0], threading.Event()

        def poll():
            while not self.nvml_stop.is_set():
                try:
                    self.nvml_sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.nvml_mask[0] |= int(get_reasons(h))
                except Exception:
                    pass
                self.nvml_stop.wait(0.002)
        self.nvml = threading.Thread(target=poll, daemon=True)
        self.nvml.start()

    def _stop_nvml(self, out):
        self.nvml_stop.set()
        self.nvml.join(timeout=2)
        if self.nvml_sm:
            out.update(sm_mhz=float(np.median(self.nvml_sm)), sm_max_mhz=self.nvml_max, samples=len(self.nvml_sm))
        out["reasons"] = sorted(nm for bit, nm in self.REASONS if self.nvml_mask[0] & bit)
        out["sampler"] = "nvml"
        return out

    def __init__(self, device_index):
        self.proc = None
        self.nvml = None
        try:
            self._start_nvml(device_index)
            return
        except Exception:
            self.nvml = None                  # fall back to an nvidia-smi child process
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
        if self.nvml is not None:
            return self._stop_nvml(out)
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
# the CPU arm: baseline/cpu_sga.py, the reference loop restated on PyTorch-CPU, on the workload's own batch
# ---------------------------------------------------------------------------------------------------
def cpu_arm(wl, steps, warmup, budget_s):
    """Times `steps` iterations of the workload's reference batch on the host cores.  The whole batch is processed every
    iteration (same config as the GPU arm); only if the projected run would exceed `budget_s` is the batch cut to its first
    k images and the rate scaled by k/n (stated in `sample`, same_config False)."""
    import sga_b200
    from baseline import cpu_sga
    if wl["variant"] != "sga":
        return None
    n_batch = wl["n"] if wl["n"] <= 24 else 7                   # Tecnick: one reference batch of 7 (configs.py:8-9)
    W = sga_b200.weights.make_synthetic(wl["c"], 0)
    x = sga_b200.synth.synthetic_images(n_batch, wl["h"], wl["w"], 0)
    probe = cpu_sga.run(x[:1], W, wl["lmbda"], 1, 1)            # 1 image: warms the CPU kernels and sizes the run
    projected = probe["seconds"][0] * n_batch * (steps + warmup)
    k = n_batch
    if projected > budget_s:
        k = max(1, int(n_batch * budget_s / projected))
    r = cpu_sga.run(x[:k], W, wl["lmbda"], steps, warmup, chunk=4)
    t = float(np.mean(r["seconds"]))
    its = (1.0 / t) * (k / float(n_batch)) * (n_batch / float(wl["n"]))
    sample = ("%d timed iterations (+%d warm-up) of %s%d images of %dx%d per iteration, autograd backward, host Adam, x "
              "re-fed every iteration (baseline/cpu_sga.py: PyTorch-CPU restatement of the reference TF1 graph, %d threads, "
              "chunks of %d images)" % (steps, warmup, "the full batch: " if k == wl["n"] else "", k, wl["h"], wl["w"],
                                        r["cores"], r["chunk"]))
    if k != wl["n"]:
        sample += "; it/s scaled to the %d-image job linearly in images" % wl["n"]
    return dict(value=its, cores=r["cores"], sample=sample, same_config=(k == wl["n"]), seconds_per_iteration=t,
                images_per_iteration=k, rd_loss_last=r["last"]["rd_loss"])


def main_reference(args, rank, wl):
    if rank != 0:
        return
    r = cpu_arm(wl, args.steps, args.warmup, budget_s=280.0)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU arm for workload %s" % args.workload}), flush=True)
        return
    v = r["value"]
    line = {
        "impl": "reference", "metric": wl["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "same_config": r["same_config"], "images_per_iteration": r["images_per_iteration"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def rank_segments(wl, rank, ws, images_override=None):
    """[(n_ref, first_image_in_batch, batch_index, global_lo, n_local)] of this rank (sga_b200.dist.rank_segments)."""
    from sga_b200 import dist as D
    if images_override:
        return [(wl["n"] if wl["n"] <= 24 else 7, 0, 0, rank * images_override, images_override)], images_override * ws
    total = wl["n"]
    if wl["n"] > 24:
        total = min(wl["n"], TECNICK_PER_GPU * ws)
    segs = D.rank_segments(total, wl["h"] * wl["w"], rank, ws)
    return [(s["n_ref"], s["first_in_batch"], s["batch"], s["lo"], s["hi"] - s["lo"]) for s in segs], total


def main_ours(args, rank, ws, local, wl):
    import torch
    import torch.distributed as dist
    import sga_b200
    from sga_b200.engine import BbEngine, SgaEngine
    from sga_b200.schedule import annealed_temperature

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    C, H, Wd, lmbda, bb = wl["c"], wl["h"], wl["w"], wl["lmbda"], wl["variant"] == "bb"
    segs, total_images = rank_segments(wl, rank, ws, args.images)
    W = sga_b200.weights.make_synthetic_bb(C, 0) if bb else sga_b200.weights.make_synthetic(C, 0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    K, Wm = args.steps, args.warmup
    temp = lambda it: annealed_temperature(it, r=1e-3, ub=0.5, scheme="exp0", t0=700)

    # one engine per segment (a contiguous image range inside ONE reference batch: n_ref and the noise offset are per batch)
    jobs = []
    all_imgs = sga_b200.synth.synthetic_images(total_images, H, Wd, 0) if total_images <= 24 else None
    for (n_ref, first, batch, g_lo, n_loc) in segs:
        if n_loc == 0:
            continue
        # the images of a batch do not depend on the sharding; large inputs are generated per segment (content is irrelevant)
        imgs = all_imgs[g_lo:g_lo + n_loc] if all_imgs is not None else sga_b200.synth.synthetic_images(n_loc, H, Wd, g_lo)
        x_host = torch.from_numpy(imgs.astype(np.float32) / 255.0).pin_memory()
        x = x_host.to(dev)
        Eng = BbEngine if bb else SgaEngine
        eng = Eng(W, n_loc, H, Wd, C, lmbda, n_ref=n_ref, operand=args.operand, seed=batch, first_image=first)
        if bb:
            y0 = eng.encode_init(x)
            zm0, zl0 = eng.hyper_init(y0)
            init = (y0, zm0, zl0)
        else:
            init = eng.encode_init(x)
        jobs.append(dict(eng=eng, x=x, x_host=x_host, init=init, n=n_loc))
    torch.cuda.synchronize()
    n_loc_total = sum(j["n"] for j in jobs)

    def new_state(j):
        return j["eng"].new_state(*[t.clone() for t in j["init"]])

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

    def run_iterations(states, first, count, rate_only=False):
        """`count` device-resident iterations of every engine of this rank (interleaved in short segments when a rank holds
        several reference batches, so that they advance together)."""
        seg = count if len(jobs) == 1 else 8
        for s0 in range(0, count, seg):
            n = min(seg, count - s0)
            for j, st in zip(jobs, states):
                if bb:
                    for k in range(n):
                        it = first + s0 + k
                        j["eng"].step(None if rate_only else j["x"], st, 1.0 if rate_only else temp(it), it,
                                      0.003 if rate_only else 0.005, rate_only=rate_only, want_stats=False)
                else:
                    j["eng"].loop(j["x"], st, [temp(first + s0 + k) for k in range(n)], first + s0, want_stats=False)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- (1) device-resident loop: `value`
    states = [new_state(j) for j in jobs]
    run_iterations(states, 0, Wm)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = jobs[0]["eng"].launch_count() if jobs else 0
    ms_total = timed(lambda: run_iterations(states, Wm, K))
    launches = (jobs[0]["eng"].launch_count() - l0) if jobs else 0
    ms_per_step = ms_total / K
    value = 1000.0 / ms_per_step

    # ---- (2) the same K iterations with one CUDA event pair around every launch (single stream, no graph, no PDL overlap)
    for j in jobs:
        j["eng"].profile(True)
    K_instr = min(K, 200)                   # the library keeps at most 16384 event pairs: ~19 launches x 200 iterations
    ms_instr = timed(lambda: run_iterations(states, Wm + K, K_instr)) / K_instr
    clocks = sampler.stop() if sampler else None
    prof = {}
    for j in jobs:
        for k, (ms, cnt) in j["eng"].profile_read().items():
            a = prof.setdefault(k, [0.0, 0])
            a[0] += ms; a[1] += cnt
        j["eng"].profile(False)

    # bits-back: the second, rate-only loop (bb_sga.py:239-261) is timed separately
    rate_loop = None
    if bb:
        st2 = []
        for j, st in zip(jobs, states):
            yt = torch.round(st["y"])
            zm, zl = j["eng"].hyper_init(yt)
            st2.append(j["eng"].new_state(yt, zm, zl))
        run_iterations(st2, 0, Wm, rate_only=True)
        ms_r = timed(lambda: run_iterations(st2, Wm, K, rate_only=True)) / K
        rate_loop = {"value": 1000.0 / ms_r, "unit": UNIT, "ms_per_step": ms_r,
                     "what": "rate-only loop over [z_mean, z_logvar] with y_tilde = round(y) (bb_sga.py:239-261): h_s fwd + "
                             "dgrad and the elementwise rate kernels, no synthesis transform"}

    # ---- (3) end to end through the public API, one iteration per call, x from pinned host memory every iteration
    states2 = [new_state(j) for j in jobs]
    copy_stream = torch.cuda.Stream()
    for j in jobs:
        j["xbuf"] = [torch.empty_like(j["x"]), torch.empty_like(j["x"])]
        j["ready"] = [torch.cuda.Event(), torch.cuda.Event()]
        j["freed"] = [torch.cuda.Event(), torch.cuda.Event()]
        j["stats_host"] = torch.empty((K + Wm, j["eng"].n_stats), dtype=torch.float32).pin_memory()
        for b in range(2):
            j["freed"][b].record(stream)

    def e2e_steps(first, count):
        for k in range(count):
            it = first + k
            b = k & 1
            for j, st in zip(jobs, states2):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(j["freed"][b])
                    j["xbuf"][b].copy_(j["x_host"], non_blocking=True)      # host -> device every iteration (sga.py:212-214)
                    j["ready"][b].record(copy_stream)
                stream.wait_event(j["ready"][b])
                if bb:
                    j["eng"].step(j["xbuf"][b], st, temp(it), it, 0.005, want_stats=True)
                else:
                    j["eng"].loop(j["xbuf"][b], st, [temp(it)], it, want_stats=True)
                j["freed"][b].record(stream)
                j["stats_host"][it].copy_(j["eng"].stats, non_blocking=True)   # device -> host read of the statistics
        stream.synchronize()

    e2e_steps(0, Wm)
    e2e_ms = timed(lambda: e2e_steps(Wm, K)) / K
    last = None
    for j in jobs:
        if not np.isfinite(j["stats_host"].numpy()).all():
            raise RuntimeError("non-finite SGA statistics inside the timed loop: the measurement is void")
        last = j["eng"].unpack_stats(j["stats_host"][Wm + K - 1].numpy())
    h2d = sum(int(j["x_host"].numel()) * 4 for j in jobs)
    d2h = sum(int(j["eng"].n_stats) * 4 for j in jobs)
    ws_bytes = sum(j["eng"].workspace_bytes for j in jobs)
    if ws > 1:
        t = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        h2d, d2h = int(t[0].item()), int(t[1].item())

    if rank != 0:
        return
    # ---- accounting (rank 0's shard; every rank runs the same kernels on its own images)
    pk = peaks()
    fl, flops_iter = {}, 0.0
    l3b = {"l3_fwd": 0.0, "l3_dgrad": 0.0}
    for j in jobs:
        f, tot = flops_per_iteration(j["n"], H, Wd, C)
        for k, v in f.items():
            fl[k] = fl.get(k, 0.0) + v
        flops_iter += tot
        for k, v in layer3_bytes(j["n"], H, Wd, C).items():
            l3b[k] += v
    fused = C in (64, 128, 192, 256)
    l3_fused = C in (64, 128, 192, 256) and not (C == 256 and os.environ.get("SGA_L3_NO_FUSE256"))
    conv_labels = {"hs0_fwd": "hs0", "hs1_fwd": "hs1", "hs2_fwd": "hs2", "hs0_dgrad": "hs0", "hs1_dgrad": "hs1",
                   "hs2_dgrad": "hs2", "gs0_fwd": "gs0", "gs1_fwd": "gs1", "gs2_fwd": "gs2", "gs0_dgrad": "gs0",
                   "gs1_dgrad": "gs1", "gs2_dgrad": "gs2"}
    # algorithmic FLOPs executed inside each launch group: the conv plus the IGDN channel mixing fused into it (forward:
    # IGDN_i after conv_i; backward: IGDN_{i-1} in the dgrad of conv_i, IGDN_2 in the layer-3 dgrad)
    fused_extra = {"gs0_fwd": "igdn0", "gs1_fwd": "igdn1", "gs2_fwd": "igdn2", "gs1_dgrad": "igdn0", "gs2_dgrad": "igdn1",
                   "l3_dgrad": "igdn2"}
    kfl = {k: fl[v] for k, v in conv_labels.items()}
    kfl["l3_fwd"] = fl["gs3"]
    kfl["l3_dgrad"] = fl["gs3"]
    if fused:
        for k, v in fused_extra.items():
            if k == "l3_dgrad" and not l3_fused:
                kfl["igdn_bwd"] = fl[v]
            else:
                kfl[k] += fl[v]
    else:
        kfl["igdn_fwd"] = sum(fl["igdn%d" % i] for i in range(3))
        kfl["igdn_bwd"] = kfl["igdn_fwd"]
    traffic = measured_traffic()
    kernels = {}
    for k, (ms, cnt) in sorted(prof.items()):
        ms_it = ms / K_instr                            # this launch group's time per iteration (all engines of the rank)
        row = {"ms_per_step": round(ms_it, 4), "launches_per_step": cnt / float(K_instr)}
        if k in ("l3_fwd", "l3_dgrad"):
            ach = l3b[k] / (ms_it * 1e-3) / 1e9
            row.update(bound="hbm", achieved=round(ach, 1), peak=pk["hbm"], unit="GB/s", frac=round(ach / pk["hbm"], 4),
                       tflops=round(kfl[k] / (ms_it * 1e-3) / 1e12, 1))
        elif k in kfl:
            ach = kfl[k] / (ms_it * 1e-3) / 1e12
            row.update(bound="tensor", achieved=round(ach, 1), peak=pk["tflops"], unit="TFLOP/s", frac=round(ach / pk["tflops"], 4))
        if traffic.get("workload") == args.workload and k in traffic.get("kernels", {}) and not args.images:
            row["traffic"] = traffic["kernels"][k] * (n_loc_total / float(wl["n"]))     # captured on the whole batch on one GPU
        kernels[k] = row
    dom = "gs2_fwd"
    dom_row = kernels.get(dom, {})
    tensor_ms = sum(v[0] for k, v in prof.items() if k in kfl) / K_instr
    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": ws, "steps": K, "warmup": Wm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if wl["n"] <= 24 else "weak",
        "vs_baseline": None, "dtype": "bf16" if args.operand == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "images_total": total_images, "images_per_gpu": n_loc_total,
                   "engines_per_gpu": len(jobs),
                   "operand": args.operand + " operands, fp32 accumulate (TMEM); latents/Adam fp32",
                   "cache": "inputs larger than L2 (%.1f GB of activations rewritten every step)" % (ws_bytes / 1e9),
                   "loop": "sga_loop: two-stream step issued from inside the library, x resident" if not bb
                           else "sga_bb_step per iteration (eager launches)",
                   "final_step": {"rd_loss": last["rd_loss"], "train_bpp": last["train_bpp"], "train_mse": last["train_mse"]}},
        "e2e": {"value": 1000.0 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor",
                     "kernel": ("%s (g_s layer 2 forward: transposed 5x5 conv + fused inverse GDN), the step's largest launch"
                                % ("k_conv_igdn_fwd<%d,2>" % C if C <= 192 else "k_conv_tc<1,256,4,2>")) if fused
                               else "k_conv_tc (g_s layer 2 forward, un-fused)",
                     "achieved": dom_row.get("achieved"), "peak": pk["tflops"], "unit": "TFLOP/s", "frac": dom_row.get("frac"),
                     "traffic": dom_row.get("traffic"),
                     "traffic_source": TRAFFIC_FILE if dom_row.get("traffic") is not None else None,
                     "peak_source": pk["source"],
                     "timed": "second pass: %d more iterations with a CUDA event pair around every launch on the launching "
                              "stream, single stream (%.3f ms/step; the value pass is un-instrumented)" % (K_instr, ms_instr),
                     "whole_step": {"gflop_per_iteration": round(flops_iter / 1e9, 1),
                                    "tflops": round(flops_iter * value / 1e12 / 1.0, 1) if ws == 1 else None,
                                    "frac_of_tensor_peak": round(flops_iter * 1000.0 / ms_per_step / 1e12 / pk["tflops"], 4),
                                    "tensor_kernels_ms": round(tensor_ms, 4)},
                     "kernels": kernels},
        "clocks": clocks,
    }
    if rate_loop:
        line["rate_loop"] = rate_loop
    if ws == 1 and not args.no_cpu:
        r = cpu_arm(wl, 2, 1, budget_s=45.0)
        if r is not None:
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                    "same_config": r["same_config"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kodak192", choices=sorted(WORKLOADS))
    ap.add_argument("--operand", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--images", type=int, default=0, help="diagnostic: images per GPU (e.g. 3 = the 8-GPU shard on one GPU); "
                                                          "not a BASELINE configuration")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, wl)
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
    main_ours(args, rank, ws, local, wl)
    if ws > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
