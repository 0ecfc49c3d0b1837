"""Multi-GPU plumbing of the flow_diffuser path: one process per GPU, the batch
is sharded by rank.  Sampling has NO exchange step (independent samples); the only collectives are the
``sync_dist=True`` scalar metrics of validation (flow_diffuser.py:281,345) and bench.py's max-over-ranks
timing.  Training's gradient all-reduce (exp_base.py:198) belongs to the backward path (next round).
Works on NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of ``range(n_items)`` for ``rank`` (first ``n % world`` ranks get one extra)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_batch(batch, rank: int, world: int):
    """Slice every tensor of a batch tuple along dim 0 for this rank."""
    n = batch[0].shape[0]
    r = shard_range(n, rank, world)
    return tuple(t[r.start:r.stop] for t in batch)


def reduce_metrics(metrics: Dict[str, float], device=None) -> Dict[str, float]:
    """Mean over ranks of a dict of scalars (Lightning's ``sync_dist=True``); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not metrics:
        return dict(metrics)
    keys = sorted(metrics)
    t = torch.tensor([float(metrics[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t /= dist.get_world_size()
    return {k: float(v) for k, v in zip(keys, t)}


def max_over_ranks(seconds: float, device=None) -> float:
    """Device-timed duration -> max over ranks (the number a multi-GPU bench must report)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
