"""Host logic of the batched eval sweep (no GPU): the global work list, its sharding, and the per-item generators that
make an item's image independent of batching and of the GPU count (BASELINE config 3)."""
import torch

from mirrorfusion_b200.sharding import shard_range
from mirrorfusion_b200.sweep import EvalSweep, item_generator


def test_item_list_is_sample_major_and_shards_partition_it():
    sw = EvalSweep.__new__(EvalSweep)          # host logic only: no engines, no CUDA
    sw.repeats = 4
    items = sw.items(5)
    assert items[:5] == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0)] and len(items) == 20
    for world in (1, 2, 3, 8):
        got = [j for r in range(world) for j in shard_range(len(items), r, world)]
        assert got == list(range(len(items)))                      # contiguous blocks, nothing lost or duplicated
        sizes = [len(shard_range(len(items), r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_item_generators_depend_on_seed_item_and_stream_only():
    a = torch.randn(4, 8, 8, generator=item_generator(7, 13, 0))
    assert torch.equal(a, torch.randn(4, 8, 8, generator=item_generator(7, 13, 0)))        # reproducible anywhere
    others = [item_generator(7, 13, 1), item_generator(7, 14, 0), item_generator(8, 13, 0)]
    assert all(not torch.equal(a, torch.randn(4, 8, 8, generator=g)) for g in others)
    seeds = {item_generator(s, i, k).initial_seed() for s in range(3) for i in range(50) for k in range(2)}
    assert len(seeds) == 3 * 50 * 2                                                          # no collisions in a sweep
