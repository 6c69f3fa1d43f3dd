"""Multi-GPU partitioning of the denoise workload: images are independent through the whole loop, so ranks take
contiguous blocks of the global image index list (the same rule as accelerate's
`PartialState.split_between_processes`, used by E/test_brushnet.py:163-168) and the only collective is the final
gather of the latents (north_star)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block of rank `rank`; the first `n_items % world` ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_list(items: Sequence, rank: int, world: int) -> List:
    r = shard_range(len(items), rank, world)
    return [items[i] for i in r]


def gather_latents(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather per-rank latents [n_local, C, H, W] (ragged over ranks) into global image order on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [len(shard_range(n_items, r, world)) for r in range(world)]
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, expected {counts[rank]}")
    mx = max(counts)
    if mx == 0:
        return local.new_zeros((0,) + tuple(local.shape[1:]))
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)
