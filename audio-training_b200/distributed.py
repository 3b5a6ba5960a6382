"""Multi-GPU layout of the front-end: one process per GPU, clips sharded across ranks, no collective on the hot
path (every 3 s clip is independent; PCEN couples frames only inside a clip).

NCCL (through torch.distributed) appears in exactly two optional places, both outside the per-clip path:
  * gather_features(): the caller wants every rank's features / predictions on one rank;
  * global_extremes(): the caller insists that PCEN's tensor-global min-max (tfpcen.py:105-110) spans all ranks.
The reference itself has no multi-GPU code at all (audiomodel.py:498-500 comments MirroredStrategy out).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from torchrun's environment; (0, 1, 0) when run alone."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init(backend=None):
    """Join the process group torchrun set up (MASTER_ADDR/PORT from the env).  No-op for a single process."""
    rank, size, local = world()
    if size > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            # device_id binds the communicator to this rank's GPU at creation (without it NCCL guesses the device at the
            # first collective and warns that a wrong guess can hang)
            dist.init_process_group(backend=backend, rank=rank, world_size=size, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=size)
    return rank, size, local


def shard_indices(n_clips, rank, world_size, mode="strided"):
    """Global clip indices owned by `rank`.  strided: clip i -> rank i mod R (keeps the synthetic
    generator sharding-invariant); block: contiguous ranges with the remainder spread over the first ranks."""
    if mode == "strided":
        return torch.arange(rank, n_clips, world_size, dtype=torch.int64)
    base, extra = divmod(n_clips, world_size)
    start = rank * base + min(rank, extra)
    return torch.arange(start, start + base + (1 if rank < extra else 0), dtype=torch.int64)


def gather_features(local, indices, n_clips, dst=0, group=None):
    """Collect per-rank features [n_local, ...] into [n_clips, ...] on rank `dst`, rows placed at their global clip
    index.  Ragged shards are padded to the largest shard for the collective.  Returns None on other ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = local.new_empty((n_clips,) + tuple(local.shape[1:]))
        out[indices.to(local.device)] = local
        return out
    size, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [len(shard_indices(n_clips, r, size)) for r in range(size)] if indices is None else None
    n_local = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    all_n = [torch.zeros_like(n_local) for _ in range(size)]
    dist.all_gather(all_n, n_local, group=group)
    all_n = [int(t.item()) for t in all_n] if counts is None else counts
    width = max(all_n)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    idx = torch.full((width,), -1, dtype=torch.int64, device=local.device)
    idx[: local.shape[0]] = indices.to(local.device)
    bufs = [torch.empty_like(pad) for _ in range(size)] if rank == dst else None
    ibufs = [torch.empty_like(idx) for _ in range(size)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    dist.gather(idx, ibufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = local.new_empty((n_clips,) + tuple(local.shape[1:]))
    for b, i, n in zip(bufs, ibufs, all_n):
        out[i[:n]] = b[:n]
    return out


def global_extremes(local_min, local_max, group=None):
    """All-reduce a (min, max) pair across ranks (two floats over NVLink): PCEN norm_scope="tensor" spanning ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_min, local_max
    lo = local_min.clone()
    hi = local_max.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return lo, hi


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (bench timing: the slowest rank defines the step)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
