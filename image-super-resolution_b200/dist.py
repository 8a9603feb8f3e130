"""Multi-GPU plumbing for the inference path: one process per GPU, images are independent
units sharded round-robin with NO data-path collective (the reference's own
extractors shard the same way: scripts/extract_test_tta_cache.py:196 ``files[rank::n]``).
torch.distributed is used only for the barrier / max-over-ranks timing reduction.
"""
from __future__ import annotations

import os
from typing import List, Sequence, TypeVar

import torch
import torch.distributed as dist

T = TypeVar("T")


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Indices of the work items (images / TTA variants) this rank processes."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def shard(items: Sequence[T], rank: int, world: int) -> List[T]:
    return [items[i] for i in shard_indices(len(items), rank, world)]


def max_over_ranks(value: float, device=None) -> float:
    """Job time = slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
