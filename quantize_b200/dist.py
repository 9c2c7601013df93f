"""Batch-sharded multi-GPU driver for the hot path.

The path shards by images: every rank holds the replicated weights / scales / quantizer parameters and processes its
own slice of the batch; there is NO collective on the data path.  torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU tests) is used only for: the barrier around timed regions, the max over ranks of the device time, and
gathering per-rank outputs for verification.
"""
import os

import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def init_from_env(backend=None):
    """one process per GPU; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the launcher (torchrun)"""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return world()


def shard_range(n_items, rank=None, world_size=None):
    """contiguous [lo, hi) slice of a global batch for `rank`; sizes differ by at most one (ragged batches)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x):
    lo, hi = shard_range(x.shape[0])
    return x[lo:hi].contiguous()


def gather_batch(local, n_items):
    """all ranks' slices concatenated in rank order (verification only).  Ragged shards are padded for the collective."""
    r, w = world()
    if w == 1:
        return local
    sizes = [shard_range(n_items, i, w) for i in range(w)]
    longest = max(hi - lo for lo, hi in sizes)
    padded = local.new_zeros((longest,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(w)]
    dist.all_gather(out, padded)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


def max_over_ranks(value, device="cpu"):
    """device time of a multi-GPU run = the slowest rank's"""
    r, w = world()
    if w == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    r, w = world()
    if w > 1:
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
