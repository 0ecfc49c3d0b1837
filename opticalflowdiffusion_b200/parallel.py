"""Multi-GPU plumbing of the flow_diffuser path: one process per GPU, the batch
is sharded by rank.  Sampling has NO exchange step (independent samples); the only collectives are the
``sync_dist=True`` scalar metrics of validation (flow_diffuser.py:281,345) and bench.py's max-over-ranks
timing.  Training's gradient exchange (exp_base.py:198) is ``optim.GradSync`` (bucketed, overlapped with the backward);
``init_distributed`` / ``sync_module_from_rank0`` are what DDPStrategy does at start-up: one process group, identical
replicas.  Works on NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_distributed(device=None, backend: str = "") -> bool:
    """Join the job's process group when launched with WORLD_SIZE > 1 (torchrun).  Idempotent.  Returns True when a
    group of more than one rank is active.  Backend: ``FD_DIST_BACKEND`` or NCCL on CUDA devices, gloo otherwise."""
    _, world, _ = env_rank()
    if world <= 1:
        return False
    if not dist.is_initialized():
        backend = backend or os.environ.get("FD_DIST_BACKEND", "") or \
            ("nccl" if device is not None and torch.device(device).type == "cuda" else "gloo")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = torch.device(device)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend, **kw)
    return dist.get_world_size() > 1


@torch.no_grad()
def sync_module_from_rank0(module: torch.nn.Module) -> int:
    """Broadcast every parameter and buffer from rank 0 (what DDP does at construction): replicas start identical
    whatever their local RNG state was.  Returns the number of tensors sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    n = 0
    seen = set()
    for t in list(module.parameters()) + list(module.buffers()):
        if t.data_ptr() in seen:            # aliased registrations (unet.* / _model.* / model.*) share storage
            continue
        seen.add(t.data_ptr())
        dist.broadcast(t.data, src=0)
        n += 1
    return n


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
