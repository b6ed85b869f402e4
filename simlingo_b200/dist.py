"""Batch-sharded data parallelism helpers (one process per GPU, torch.distributed for the plumbing).

The hot path shards by independent samples (reference: Lightning DDP / ZeRO-2 are both batch-sharded DP,
``train_simlingo_seed1.sh:27``, ``config.py:299``): inference needs no data-path collective; training has exactly
one exchange step, the gradient all-reduce (``simlingo_b200.training``)."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of ``n_items`` independent samples for ``rank`` (first ranks get the
    remainder), e.g. 64 frames over 8 ranks -> 8 each."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Step time of the job = slowest rank (device-timed milliseconds are reduced with MAX)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_predictions(local: torch.Tensor, sizes: Sequence[int]) -> List[torch.Tensor] | None:
    """Optional final gather of per-rank outputs (e.g. route [b_r, 20, 2]) to rank 0; KBs, off the hot path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = max(sizes)  # gather needs equal shapes: pad the ragged last shard, trim after
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0)
    return [b[: sizes[r]] for r, b in enumerate(bufs)] if rank == 0 else None
