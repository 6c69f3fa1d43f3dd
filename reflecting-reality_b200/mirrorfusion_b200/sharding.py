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


def grad_buckets(numel: int, bucket_elems: int) -> List[range]:
    """Contiguous element ranges of the flat gradient buffer, one per collective call."""
    if bucket_elems <= 0:
        raise ValueError("bucket_elems must be positive")
    return [range(s, min(s + bucket_elems, numel)) for s in range(0, numel, bucket_elems)]


def allreduce_flat_grads(flat_grad: torch.Tensor, group=None, bucket_elems: int = 64 << 20, async_op: bool = False):
    """SUM all-reduce of the flat BrushNet gradient buffer in large contiguous buckets (config 4: 618.8 M fp32 gradients =
    10 calls of 256 MB; NCCL over NVLink picks NVLS / ring itself).  The 1 / world mean is NOT applied here: it is the
    `grad_scale` of the fused clip + AdamW kernel, so the gradients are read once.  The reference gets the same reduction
    from DDP inside `accelerator.backward` (E/train_brushnet_mirror.py:1459).  Returns the work handles if async_op."""
    import torch.distributed as dist
    works = []
    for r in grad_buckets(flat_grad.numel(), bucket_elems):
        w = dist.all_reduce(flat_grad[r.start:r.stop], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def gather_loss(loss: torch.Tensor, group=None) -> float:
    """Mean over ranks of the per-rank loss, for logging (`accelerator.gather(loss.repeat(bs)).mean()`,
    E/train_brushnet_mirror.py:1453-1456)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [torch.empty_like(loss) for _ in range(world)]
    dist.all_gather(out, loss, group=group)
    return float(torch.stack(out).mean().item())
