"""The `compress` command of the reference's sga.py, on the B200 engine.

Same flags (tf_boilerplate.parse_args, tf_boilerplate.py:91-204), same checkpoint directory layout
(<checkpoint_dir>/<runname>/, sga.py:180-182), same progress lines (sga.py:226-233), same output file
rd-sga-lmbda=%g+<runname>-input=<file>.npz with the 7 per-image fields (sga.py:183, 258-269) and the
`Avg <field>` summary (sga.py:293-295).  Deviation, documented: --sga_its is honoured (the reference parses it but
hard-codes 2000, sga.py:191-192)."""
import argparse
import os
import sys

import numpy as np

SCRIPT_NAME = "sga"


def parse_args(argv, bb=False):
    try:
        from absl.flags import argparse_flags
        parser = argparse_flags.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    except Exception:  # absl is optional
        parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--verbose", "-V", action="store_true",
                        help="Report bitrate and distortion when training or compressing.")
    parser.add_argument("--num_filters", type=int, default=-1, help="Number of filters in the latents.")
    parser.add_argument("--num_hfilters", type=int, default=-1, help="Number of filters in the hyper latents.")
    parser.add_argument("--checkpoint_dir", default="./checkpoints", help="Directory where to save/load model checkpoints.")
    sub = parser.add_subparsers(title="commands", dest="command", help="Only 'compress' is supported (sga.py:303).")
    c = sub.add_parser("compress", formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                       description="Runs SGA iterative inference on an image or a .npy batch and writes RD statistics.")
    c.add_argument("--results_dir", default="./results", help="Directory for storing compression stats/results.")
    c.add_argument("--lambda", type=float, default=-1, dest="lmbda", help="Lambda for rate-distortion tradeoff.")
    c.add_argument("--sga_its", type=int, default=2000, help="Number of SGA iterations.")
    c.add_argument("--annealing_rate", type=float, default=1e-3, help="Annealing rate for SGA.")
    c.add_argument("--t0", type=int, default=700, help="Number of 'soft-quantization' iterations before annealing.")
    if bb:
        c.add_argument("--rate_its", type=int, default=2000,
                       help="(new) iterations of the second, rate-only loop (bb_sga.py hard-codes r_opt_its = 2000).")
    if not bb:
        c.add_argument("--save_opt_record", action="store_true",
                       help="(new flag for configs.save_opt_record) log every 10 iterations and write opt-*.npz (sga.py:188-189, 271-278).")
        c.add_argument("--save_reconstruction", action="store_true",
                       help="(new flag for sga.py's save_reconstruction) write recon-*.png of a single-image input (sga.py:280-291).")
    c.add_argument("--operand", choices=["bf16", "f32"], default="bf16",
                   help="(new) arithmetic of the (de)conv stacks inside the loop; init and final eval are always f32.")
    c.add_argument("runname", help="Model name identifier, like 'mbt2018-num_filters=192-lmbda=0.01'")
    c.add_argument("input_file", help="Input filename (image, or .npy uint8 [N,H,W,3]).")
    c.add_argument("output_file", nargs="?", help="Output filename (optional; the reference never writes it -- here: the "
                                                  "range-coded bitstream of round(y), round(z), single-process runs).")
    args = parser.parse_args(argv[1:])
    if args.command is None:
        parser.print_usage()
        sys.exit(2)
    return args


def load_input(path):
    if path.endswith(".npy"):
        X = np.load(path)                                     # sga.py:41-43
    else:
        from PIL import Image
        X = np.asarray(Image.open(path).convert("RGB"))[None, ...]    # sga.py:45-48
    if X.ndim != 4 or X.shape[-1] != 3:
        raise ValueError("expected images of shape [N,H,W,3], got %s" % (X.shape,))
    return X


def compress(args, out=print):
    import torch
    from . import dist as D
    from . import weights as Wm
    from .engine import EVAL_FIELDS, SgaEngine
    from .msssim import msssim_per_image
    from .schedule import annealed_temperature

    rank, ws, local = D.init_process_group()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    say = out if rank == 0 else (lambda *a, **k: None)
    X = load_input(args.input_file)
    num_images, H, Wd = int(X.shape[0]), int(X.shape[1]), int(X.shape[2])
    if args.lmbda < 0:                                        # sga.py:157-159
        args.lmbda = float(args.runname.split("lmbda=")[1].split("-")[0])
        say("Defaulting lmbda (mse coefficient) to %g as used in model training." % args.lmbda)
    W = Wm.load_checkpoint(args.checkpoint_dir, args.runname, args.num_filters)
    batches = D.reference_batches(num_images, H * Wd)
    # This rank's images: a contiguous global range cut at the reference's batch boundaries.  Every segment is one engine
    # with its batch's n_ref; the Philox key is the batch index and the counter offset the position inside the batch, so
    # every image of the input gets its own noise whatever the sharding (the reference's TF stream also keeps advancing
    # from batch to batch).  All segments advance together, so a many-batch input keeps every GPU busy with all of its
    # images at once instead of walking the batches one after the other.
    segs = D.rank_segments(num_images, H * Wd, rank, ws)
    jobs = []
    for sg in segs:
        eng = SgaEngine(W, sg["hi"] - sg["lo"], H, Wd, args.num_filters, args.lmbda, n_ref=sg["n_ref"], operand=args.operand,
                        seed=sg["batch"], first_image=sg["first_in_batch"])
        x = torch.from_numpy(X[sg["lo"]:sg["hi"]].astype(np.float32) / 255.0).cuda()                 # sga.py:52-53
        y, z = eng.encode_init(x)                                                                    # sga.py:207
        jobs.append(dict(seg=sg, eng=eng, x=x, st=eng.new_state(y, z)))
    its = args.sga_its
    temps = [annealed_temperature(it, r=args.annealing_rate, ub=0.5, scheme="exp0", t0=args.t0) for it in range(its)]
    from . import configs
    save_opt_record = bool(getattr(args, "save_opt_record", False) or configs.save_opt_record)
    save_reconstruction = bool(getattr(args, "save_reconstruction", False) or configs.save_reconstruction)
    log_itv = 10 if save_opt_record else 100                                                         # sga.py:187-189
    log_points = [it for it in range(its) if it % log_itv == 0 or it + 1 == its]                      # sga.py:216
    opt_records = [dict(its=[], T=[], rd_loss=[], rd_loss_after_rounding=[]) for _ in batches]       # sga.py:209
    nb = len(batches)
    nf = 7 if args.verbose else 4
    lines = [[] for _ in batches]
    t_loop = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] if torch.cuda.is_available() else None
    if t_loop:
        torch.cuda.synchronize()
        t_loop[0].record()
    first = 0
    for end in log_points:
        for j in jobs:                                         # device-resident segment [first, end] of the loop (sga.py:210-215)
            j["eng"].loop(j["x"], j["st"], temps[first:end + 1], first, want_stats=True)
        acc = np.zeros((nb, nf))
        for j in jobs:
            s = j["eng"].unpack_stats(j["eng"].stats.cpu().numpy())
            row = [s["rd_loss"], s["train_mse"], s["train_bpp"], float(s["psnr"].sum())]
            if args.verbose:                                   # sga.py:218-230: the same statistics with hard-rounded latents
                ex = j["eng"].eval_hard(j["x"], j["st"]["y"], j["st"]["z"])
                row += [ex["rd_loss"], ex["train_bpp"], float(ex["psnr"].sum())]
            acc[j["seg"]["batch"]] += row
        tot = np.asarray(D.allreduce_scalars(acc.reshape(-1).tolist())).reshape(nb, nf)
        for b, (b_lo, b_hi) in enumerate(batches):
            line = "it=%d, T=%.3f rd_loss=%.4f mse=%.3f bpp=%.4f psnr=%.4f" % (end, temps[end], tot[b, 0], tot[b, 1], tot[b, 2],
                                                                              tot[b, 3] / (b_hi - b_lo))
            if args.verbose:
                line += "\t after rounding: rd_loss=%.4f, bpp=%.4f psnr=%.4f" % (tot[b, 4], tot[b, 5], tot[b, 6] / (b_hi - b_lo))
            if not np.isfinite(tot[b]).all():
                raise FloatingPointError("non-finite SGA statistics at iteration %d of batch %d: %s" % (end, b, line))
            lines[b].append(line)
            opt_records[b]["its"].append(end); opt_records[b]["T"].append(temps[end]); opt_records[b]["rd_loss"].append(tot[b, 0])
            if args.verbose:
                opt_records[b]["rd_loss_after_rounding"].append(tot[b, 4])                           # sga.py:230-236
        first = end + 1
    loop_ms = None
    if t_loop:
        t_loop[1].record()
        torch.cuda.synchronize()
        loop_ms = t_loop[0].elapsed_time(t_loop[1])
    for b in range(nb):                                        # the reference's output, batch after batch
        say("----RD Optimization----")                        # sga.py:206
        for line in lines[b]:
            say(line)
        say()
    rows = []
    for j in jobs:                                             # sga.py:240-247
        ev = j["eng"].eval_hard(j["x"], j["st"]["y"], j["st"]["z"])
        recon = j["eng"].reconstruction()
        ms = msssim_per_image(recon, j["x"])
        rows.append(np.stack([ev["mse"], ev["psnr"], ms, -10 * np.log10(np.maximum(1 - ms, 1e-12)), ev["est_bpp"],
                              ev["est_y_bpp"], ev["est_z_bpp"]], axis=1).astype(np.float32))
        if save_reconstruction:                               # sga.py:280-291 (single-image inputs only)
            assert num_images == 1
            j["recon_u8"] = torch.round(torch.clamp(recon[0], 0, 1) * 255).to(torch.uint8).cpu().numpy()   # utils.write_png
        if args.output_file and ws == 1:
            # (new) the bitstream of what SGA transmits, round(y) and round(z) (sga.py:240-241), with the range coder of
            # csrc/codec.cu -- the reference accepts an output_file argument but never writes it (sga.py only estimates)
            from . import codec as Cd
            cd = Cd.LatentCodec(j["eng"])
            j["streams"], _, _ = cd.compress(j["st"]["y"], j["st"]["z"], Cd.MODE_SGA)
            cd.close()
        j["eng"].close()
    rows = np.concatenate(rows, axis=0) if rows else np.zeros((0, 7), dtype=np.float32)
    full = D.gather_rows(rows, num_images, rank, ws, device="cuda" if torch.cuda.is_available() else "cpu")
    if loop_ms is not None and its > 0:
        loop_ms = max(D.allgather_floats(loop_ms))
        say("SGA loop: %d iterations of %d images (%d reference batch%s) on %d GPU%s in %.2f s = %.1f it/s"
            % (its, num_images, nb, "" if nb == 1 else "es", ws, "" if ws == 1 else "s", loop_ms / 1000.0,
               its / (loop_ms / 1000.0)))
    if rank != 0:
        return None
    results = {k: np.asarray(full[:, i]) for i, k in enumerate(EVAL_FIELDS)}
    input_file = os.path.basename(args.input_file)
    trained_script_name = args.runname.split("-")[0]
    save_file = "rd-%s-input=%s.npz" % (args.runname, input_file)
    if SCRIPT_NAME != trained_script_name:                    # sga.py:263-268
        save_file = "rd-%s-lmbda=%g+%s-input=%s.npz" % (SCRIPT_NAME, args.lmbda, args.runname, input_file)
    if args.output_file and ws == 1:
        from . import codec as Cd
        streams = [st_ for j in jobs for st_ in j["streams"]]
        blob = Cd.pack(streams, Cd.MODE_SGA, args.num_filters, H, Wd)
        with open(args.output_file, "wb") as f:
            f.write(blob)
        results["actual_bpp"] = np.asarray([(len(zs) + len(ys)) * 8.0 / (H * Wd) for zs, ys in streams])
        say("Wrote %s (%d bytes); Avg actual_bpp: %.4f (estimated %.4f)" % (args.output_file, len(blob), results["actual_bpp"].mean(),
                                                                        results["est_bpp"].mean()))
    elif args.output_file:
        say("output_file is only written by single-process runs")
    np.savez(os.path.join(args.results_dir, save_file), **results)
    if save_opt_record:                                       # sga.py:271-278 (the record of the last batch, as in the reference)
        opt_file = "opt-%s-input=%s.npz" % (args.runname, input_file)
        if SCRIPT_NAME != trained_script_name:
            opt_file = "opt-%s-lmbda=%g+%s-input=%s.npz" % (SCRIPT_NAME, args.lmbda, args.runname, input_file)
        np.savez(os.path.join(args.results_dir, opt_file), **{k: np.asarray(v) for k, v in opt_records[-1].items()})
    if save_reconstruction and jobs and "recon_u8" in jobs[0]:
        png = "recon-%s-input=%s.png" % (args.runname, input_file)
        if SCRIPT_NAME != trained_script_name:
            png = "recon-%s-lmbda=%g-rd_opt_its=%d+%s-input=%s.png" % (SCRIPT_NAME, args.lmbda, its, args.runname, input_file)
        png = os.path.join(args.results_dir, png)
        say("Saving image reconstruction to ", png)
        from PIL import Image
        Image.fromarray(jobs[0]["recon_u8"]).save(png)
    for k in EVAL_FIELDS:
        say("Avg {}: {:0.4f}".format(k, results[k].mean()))   # sga.py:293-295
    return results


def compress_bb(args, out=print):
    """The `compress` command of the reference's bb_sga.py (bits-back variant): R-D loop over [y, z_mean, z_logvar]
    (bb_sga.py:201-233), rate loop over [z_mean, z_logvar] with y_tilde = round(y) (:235-256), evaluation (:271-275),
    output rd-bb_sga-lmbda=%g+<runname>-input=<file>.npz with the 8 fields of bb_sga.py:176."""
    import torch
    from . import dist as D
    from . import weights as Wm
    from .engine import BB_EVAL_FIELDS, BbEngine
    from .msssim import msssim_per_image

    rank, ws, local = D.init_process_group()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    say = out if rank == 0 else (lambda *a, **k: None)
    X = load_input(args.input_file)
    num_images, H, Wd = int(X.shape[0]), int(X.shape[1]), int(X.shape[2])
    if args.lmbda < 0:                                        # bb_sga.py:150-152
        args.lmbda = float(args.runname.split("lmbda=")[1].split("-")[0])
        say("Defaulting lmbda (mse coefficient) to %g as used in model training." % args.lmbda)
    W = Wm.load_checkpoint(args.checkpoint_dir, args.runname, args.num_filters, bb=True)
    results = {k: [] for k in BB_EVAL_FIELDS}
    log_points = lambda its: [it for it in range(its) if it % 100 == 0 or it + 1 == its]
    for b_idx, (b_lo, b_hi) in enumerate(D.reference_batches(num_images, H * Wd)):
        n_ref = b_hi - b_lo
        lo, hi = D.shard_range(n_ref, rank, ws)
        n_loc = hi - lo
        rows = np.zeros((0, 8), dtype=np.float32)
        say("----RD Optimization----")                        # bb_sga.py:202

        def log(it, T, s, extra=None):                        # bb_sga.py:216-232
            tot = D.allreduce_scalars([s["rd_loss"], s["train_mse"], s["train_bpp"], float(s["psnr"].sum())]
                                      + ([extra["rd_loss"], extra["train_bpp"], float(extra["psnr"].sum())] if extra else []))
            line = "it=%d, T=%.3f rd_loss=%.4f mse=%.3f bpp=%.4f psnr=%.4f" % (it, T, tot[0], tot[1], tot[2], tot[3] / n_ref)
            if extra:
                line += "\t after rounding: rd_loss=%.4f, bpp=%.4f psnr=%.4f" % (tot[4], tot[5], tot[6] / n_ref)
            say(line)

        def log_rate(it, s):                                  # bb_sga.py:254-255
            if it == 0:
                say()
                say("----Rate Optimization----")
            say("it=", it, "\trate=", D.allreduce_scalars([s["train_bpp"]])[0])

        if n_loc > 0:
            # Philox key = batch index: every image of the input gets its own noise (the reference's stream keeps advancing)
            eng = BbEngine(W, n_loc, H, Wd, args.num_filters, args.lmbda, n_ref=n_ref, operand=args.operand, seed=b_idx,
                           first_image=lo)
            x = torch.from_numpy(X[b_lo + lo:b_lo + hi].astype(np.float32) / 255.0).cuda()
            y_tilde, zm, zl, ev = eng.run(x, its=args.sga_its, r_its=args.rate_its, annealing_rate=args.annealing_rate,
                                          t0=args.t0, log=log, log_rate=log_rate, verbose=args.verbose)
            ms = msssim_per_image(eng.reconstruction(), x)
            rows = np.stack([ev["mse"], ev["psnr"], ms, -10 * np.log10(np.maximum(1 - ms, 1e-12)), ev["est_bpp"],
                             ev["est_y_bpp"], ev["est_z_bpp"], ev["est_bpp_back"]], axis=1).astype(np.float32)
            eng.close()
        elif ws > 1:                                          # ranks without images still join the log all-reduces
            for _ in log_points(args.sga_its):
                D.allreduce_scalars([0.0] * (7 if args.verbose else 4))
            for _ in log_points(args.rate_its):
                D.allreduce_scalars([0.0])
        say()
        full = D.gather_rows(rows, n_ref, rank, ws, device="cuda" if torch.cuda.is_available() else "cpu")
        for i, k in enumerate(BB_EVAL_FIELDS):
            results[k] += full[:, i].tolist()
    if rank != 0:
        return None
    results = {k: np.asarray(v) for k, v in results.items()}
    input_file = os.path.basename(args.input_file)
    save_file = "rd-%s-input=%s.npz" % (args.runname, input_file)
    if "bb_sga" != args.runname.split("-")[0]:                # bb_sga.py:283-287
        save_file = "rd-%s-lmbda=%g+%s-input=%s.npz" % ("bb_sga", args.lmbda, args.runname, input_file)
    np.savez(os.path.join(args.results_dir, save_file), **results)
    for k in BB_EVAL_FIELDS:
        say("Avg {}: {:0.4f}".format(k, results[k].mean()))   # bb_sga.py:290-292
    return results


def parse_args_mbt2018(argv):
    """tf_boilerplate.parse_args as mbt2018.py uses it: compress and decompress (training is out of scope)."""
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--verbose", "-V", action="store_true")
    parser.add_argument("--num_filters", type=int, default=-1, help="Number of filters in the latents.")
    parser.add_argument("--num_hfilters", type=int, default=-1)
    parser.add_argument("--checkpoint_dir", default="./checkpoints", help="Directory where to save/load model checkpoints.")
    sub = parser.add_subparsers(title="commands", dest="command")
    c = sub.add_parser("compress", formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                       description="Reads an image / .npy batch, compresses it with the base model, writes RD statistics and, "
                                   "with an output file, the bitstream.")
    c.add_argument("--results_dir", default="./results")
    c.add_argument("runname")
    c.add_argument("input_file")
    c.add_argument("output_file", nargs="?", help="bitstream file (the reference writes <input>.tfci when "
                                                  "configs.write_tfci_for_eval is set)")
    d = sub.add_parser("decompress", formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                       description="Reads a bitstream file, reconstructs the image(s), writes a PNG (one image) or .npy.")
    d.add_argument("runname")
    d.add_argument("input_file")
    d.add_argument("output_file", nargs="?")
    args = parser.parse_args(argv[1:])
    if args.command is None:
        parser.print_usage()
        sys.exit(2)
    return args


def compress_mbt2018(args, out=print):
    """The base model's one-shot compression (mbt2018.py:131-245): y = g_a(x), z = h_a(y), z_hat = round(z), y_hat =
    round(y - mu) + mu, range-coded strings, the 7 eval fields plus batch_actual_bpp / batch_sizes /
    avg_batch_actual_bpp, output rd-<runname>-file=<input>.npz."""
    import torch
    from . import codec as Cd
    from . import dist as D
    from . import weights as Wm
    from .engine import EVAL_FIELDS, SgaEngine
    from .msssim import msssim_per_image
    if D.world()[0] != 0:                      # a one-shot forward pass: single process (extra torchrun ranks do nothing)
        return None
    X = load_input(args.input_file)
    num_images, H, Wd = int(X.shape[0]), int(X.shape[1]), int(X.shape[2])
    num_pixels = H * Wd
    W = Wm.load_checkpoint(args.checkpoint_dir, args.runname, args.num_filters)
    results = {k: [] for k in EVAL_FIELDS}
    batch_actual_bpp, batch_sizes, blobs = [], [], []
    for (lo, hi) in D.reference_batches(num_images, num_pixels):
        n = hi - lo
        eng = SgaEngine(W, n, H, Wd, args.num_filters, 0.01, operand="f32")
        cd = Cd.LatentCodec(eng)
        x = torch.from_numpy(X[lo:hi].astype(np.float32) / 255.0).cuda()
        y, z = eng.encode_init(x)
        streams, y_hat, z_hat = cd.compress(y, z, Cd.MODE_MEAN_CENTRED)                    # mbt2018.py:84-85
        ev = eng.eval_tilde(x, y_hat, z_hat)
        ms = msssim_per_image(eng.reconstruction(), x)
        rows = [ev["mse"], ev["psnr"], ms, -10 * np.log10(np.maximum(1 - ms, 1e-12)), ev["est_bpp"], ev["est_y_bpp"],
                ev["est_z_bpp"]]
        for k, v in zip(EVAL_FIELDS, rows):
            results[k] += np.asarray(v).tolist()
        packed = Cd.pack(streams, Cd.MODE_MEAN_CENTRED, args.num_filters, H, Wd)            # mbt2018.py:211-214
        blobs.append(packed)
        batch_actual_bpp.append(len(packed) * 8 / num_pixels)                              # mbt2018.py:220-221
        batch_sizes.append(n)
        cd.close()
        eng.close()
    if args.output_file:
        with open(args.output_file, "wb") as f:                                            # mbt2018.py:215-217
            f.write(blobs[0] if len(blobs) == 1 else b"".join(blobs))
    results = {k: np.asarray(v) for k, v in results.items()}
    results["batch_actual_bpp"] = np.asarray(batch_actual_bpp)
    results["batch_sizes"] = np.asarray(batch_sizes)
    results["avg_batch_actual_bpp"] = np.sum(batch_actual_bpp) / np.sum(batch_sizes)       # mbt2018.py:233
    np.savez(os.path.join(args.results_dir, "rd-%s-file=%s.npz" % (args.runname, os.path.basename(args.input_file))), **results)
    for k in EVAL_FIELDS + ["avg_batch_actual_bpp"]:
        out("Avg {}: {:0.4f}".format(k, np.mean(results[k])))                             # mbt2018.py:243-245
    return results


def decompress_mbt2018(args, out=print):
    """mbt2018.py:248-295: strings -> z_hat -> (mu, sigma) -> y_hat -> x_hat, written as PNG (one image) or .npy."""
    import torch
    from . import codec as Cd
    from . import dist as D
    from . import weights as Wm
    from .engine import SgaEngine
    if D.world()[0] != 0:
        return None
    with open(args.input_file, "rb") as f:
        data = f.read()
    streams, mode, C, H, Wd = Cd.unpack(data)
    if args.num_filters > 0 and args.num_filters != C:
        raise ValueError("bitstream was written with num_filters=%d, --num_filters says %d" % (C, args.num_filters))
    W = Wm.load_checkpoint(args.checkpoint_dir, args.runname, C)
    eng = SgaEngine(W, len(streams), H, Wd, C, 0.01, operand="f32")
    cd = Cd.LatentCodec(eng)
    y_hat, z_hat = cd.decompress(streams, mode)
    eng.eval_tilde(None, y_hat, z_hat)
    xt = eng.reconstruction()
    img = torch.round(torch.clamp(xt, 0, 1) * 255).to(torch.uint8).cpu().numpy()           # utils.write_png: saturate + round
    dst = args.output_file or (args.input_file + (".png" if len(streams) == 1 else ".npy"))
    if dst.endswith(".npy") or len(streams) != 1:
        np.save(dst, img)
    else:
        from PIL import Image
        Image.fromarray(img[0]).save(dst)
    cd.close()
    eng.close()
    out("wrote %s" % dst)
    return img


def main_mbt2018(argv=None):
    args = parse_args_mbt2018(sys.argv if argv is None else argv)
    if args.command == "compress":
        compress_mbt2018(args)
    elif args.command == "decompress":
        decompress_mbt2018(args)
    else:
        raise SystemExit("only compress / decompress are supported (training is out of scope)")


def main(argv=None):
    args = parse_args(sys.argv if argv is None else argv)
    assert args.command == "compress", "Only compression is supported."     # sga.py:303
    compress(args)


def main_bb(argv=None):
    args = parse_args(sys.argv if argv is None else argv, bb=True)
    assert args.command == "compress", "Only compression is supported."     # bb_sga.py:300
    compress_bb(args)
