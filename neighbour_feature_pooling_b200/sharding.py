"""Batch sharding of the NFP path across the GPUs of one box.

The operator is independent per image, so multi-GPU operation is one
process per GPU, each owning a contiguous slice of the batch, with NO collective on the data path.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only to agree on timings and to
gather scalar metrics.  The one exception is the reference's batch-coupled ``scs`` measure
(nfp.py:363-374), which cannot be sharded without changing its result -- ``check_shardable`` rejects it.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_from_env(backend: str, device=None):
    """Join the job torchrun described in the environment; no-op for a single process."""
    rank, local, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced ``[lo, hi)`` slice of ``n`` maps for ``rank`` (first ``n % world`` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def check_shardable(measure: str):
    if measure.lower() in ("scs", "sharpened_cosine"):
        raise RuntimeError("the reference's 'scs' measure averages over the batch (nfp.py:363-374): its result "
                           "changes if the batch is sharded; run it on one rank")


def max_over_ranks(value: float, device="cpu") -> float:
    """A timing is the slowest rank's."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_rows(local: torch.Tensor) -> torch.Tensor:
    """Concatenate per-rank result rows (e.g. pooled descriptors / metrics) in rank order; ragged allowed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
