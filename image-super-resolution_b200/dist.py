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


TILE_GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}


def job_schedule(n_items: int, world: int):
    """Partition of an n_items-image inference job over ``world`` ranks (BASELINE configs[2]: 100 images over 1/2/4/8 GPUs).

    Whole images go round-robin, ``items[rank::world]``, exactly as the reference's extractors shard
    (scripts/extract_test_tta_cache.py:196), for as many full rounds as there are.  The last ``t = n_items % world`` images
    would leave ``world - t`` ranks idle for a whole image time; when ``world`` is a multiple of ``t`` each of them is
    instead split into ``world / t`` halo tiles (serving.fuse_tiled) over a group of consecutive ranks (100 images on 8 GPUs:
    12 rounds + 4 images x 2 tiles).  Returns ``(whole, tail)``: ``whole[r]`` = image indices of rank r; ``tail`` = list of
    ``(image, ranks, grid)``."""
    if world < 1 or n_items < 0:
        raise ValueError("job_schedule: bad arguments")
    full = n_items // world * world
    whole = [list(range(r, full, world)) for r in range(world)]
    rest = list(range(full, n_items))
    tail = []
    if rest:
        t = len(rest)
        if world % t == 0 and world // t > 1:
            g = world // t
            for j, img in enumerate(rest):
                tail.append((img, list(range(j * g, (j + 1) * g)), TILE_GRIDS.get(g, (1, g))))
        else:
            for j, img in enumerate(rest):
                whole[j % world].append(img)
    return whole, tail
