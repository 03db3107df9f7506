"""Batch segmentation and multi-GPU sharding.

The reference processes the input in batches of round(1e7 / pixels) images (configs.py:5-9, sga.py:55-57) and both
loss terms are means over that batch (sga.py:147,150), so every gradient carries 1/n_ref.  Images are independent
otherwise: each rank takes a contiguous range of every reference batch, runs the loop with inv-batch scale 1/n_ref and
no data-path collective; per-image result rows are gathered once after the loop (NCCL over NVLink on the box, gloo in
the CPU tests).  The log lines need the batch-wide scalars, which is one tiny all-reduce per log point."""
import numpy as np

EVAL_BATCH_NUM_PIXELS = 1e7   # configs.py:5


def eval_batch_size(num_pixels_per_image):
    """configs.get_eval_batch_size (configs.py:8-9).  The pixel budget is sga_b200.configs.eval_batch_num_pixels; the
    environment variable SGA_EVAL_BATCH_NUM_PIXELS overrides it (tests use it to get several reference batches out of
    small images)."""
    import os
    from . import configs
    budget = float(os.environ.get("SGA_EVAL_BATCH_NUM_PIXELS", configs.eval_batch_num_pixels))
    return int(round(budget / num_pixels_per_image))


def reference_batches(num_images, num_pixels_per_image):
    """[(lo, hi)) image ranges of the reference's tf.data batches (sga.py:56-57)."""
    bs = max(1, eval_batch_size(num_pixels_per_image))
    return [(lo, min(lo + bs, num_images)) for lo in range(0, num_images, bs)]


def shard_range(n, rank, world_size):
    """Contiguous, balanced [lo, hi) of `n` items for `rank`."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_segments(num_images, num_pixels_per_image, rank, world_size):
    """Sharding ACROSS reference batches: the rank takes the contiguous global image range shard_range(num_images, ...) and
    cuts it at the reference's batch boundaries.  Every segment lies inside one reference batch, so it runs as one engine
    with that batch's n_ref (the 1/n_ref of every gradient, sga.py:147,150) and its own offset into the batch's noise
    stream; all segments of a rank advance together.  A 100-image Tecnick input (15 batches of 7, 7, ..., 2) then keeps
    12-13 images on each of 8 GPUs busy at once instead of 15 sequential passes with at most one image per GPU.
    Returns [dict(batch, n_ref, lo, hi, first_in_batch)] with global [lo, hi)."""
    lo, hi = shard_range(num_images, rank, world_size)
    out = []
    for b, (b_lo, b_hi) in enumerate(reference_batches(num_images, num_pixels_per_image)):
        s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
        if s_lo < s_hi:
            out.append(dict(batch=b, n_ref=b_hi - b_lo, lo=s_lo, hi=s_hi, first_in_batch=s_lo - b_lo))
    return out


def world():
    import os
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """One process per GPU; rendezvous from the torchrun environment."""
    import torch
    import torch.distributed as dist
    rank, ws, local = world()
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, ws, local


def gather_rows(rows, n_total, rank, world_size, device="cuda"):
    """rows: float32 [n_local, k] of this rank's shard (shard_range order). Returns [n_total, k] on rank 0
    (and on every rank: it is an all-gather of <= 8 floats per image)."""
    import torch
    import torch.distributed as dist
    rows = np.asarray(rows, dtype=np.float32)
    if world_size == 1:
        return rows
    k = rows.shape[1]
    cap = -(-n_total // world_size)
    buf = torch.zeros((cap, k), dtype=torch.float32, device=device)
    if rows.shape[0]:
        buf[:rows.shape[0]] = torch.from_numpy(rows).to(device)
    out = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world_size):
        lo, hi = shard_range(n_total, r, world_size)
        parts.append(out[r][:hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0)


def allreduce_scalars(vals, device="cuda"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in vals]
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return [float(v) for v in t.cpu()]


def allgather_floats(v, device="cuda"):
    """One float per rank -> list over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v)]
    if dist.get_backend() == "gloo":
        device = "cpu"
    t = torch.tensor([v], dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]
