"""Host logic of the fine-tune glue (no GPU): flat parameter layout, AdamW scalars, min-SNR weights, gradient buckets and the
world-size-2 gloo run of the flat gradient all-reduce + loss gather."""
import math
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from mirrorfusion_b200 import sharding
from mirrorfusion_b200.train import B200AdamW, FlatParams, NoiseSchedule, flat_layout
from oracle import train_oracle as T


def test_flat_layout_alignment_and_views():
    shapes = {"a.weight": (5, 3), "a.bias": (5,), "s": (), "b.weight": (8, 4, 3, 3)}
    table, total = flat_layout(shapes)
    assert all(off % 8 == 0 for off, _ in table.values())        # 16 B in the bf16 copy: TMA tensor-map bases
    assert table["a.weight"] == (0, 15) and table["a.bias"] == (16, 5) and table["s"] == (24, 1) and table["b.weight"][0] == 32
    assert total == 32 + 288
    fp = FlatParams(shapes, "cpu", with_bf16=False)
    fp.p("a.bias").fill_(2.0)
    assert fp.param[16:21].eq(2).all() and fp.param[21:24].eq(0).all()
    assert fp.g("b.weight").shape == (8, 4, 3, 3) and fp.g("b.weight").data_ptr() == fp.grad[32:].data_ptr()


def test_adamw_scalars():
    fp = FlatParams({"w": (4,)}, "cpu", with_bf16=False)
    opt = B200AdamW(fp, lr=5e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    h = opt.hyper(3, grad_scale=0.125)
    assert np.allclose(h, [5e-6, 0.9, 0.999, 1e-8, 1e-2, 1 - 0.9 ** 3, math.sqrt(1 - 0.999 ** 3), 0.125, 0.1, 0.001, 1 - 5e-8,
                           5e-6 / (1 - 0.9 ** 3)], rtol=1e-12)
    assert abs(h[9] - 0.001) < 1e-15      # NOT 1 - float32(0.999)
    opt.param_groups[0]["lr"] = 1e-4       # what an lr scheduler does
    assert opt.hyper(1)[0] == 1e-4


def test_snr_weights_match_oracle_and_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_glue.npz"))
    ns = NoiseSchedule("cpu")
    assert np.array_equal(ns.acp_host, g["alphas_cumprod"])
    t = torch.from_numpy(g["t"])
    np.testing.assert_allclose(ns.snr_weights(t, 5.0), g["w_snr5"], rtol=1e-6)
    np.testing.assert_allclose(ns.snr_weights(t, 5.0), T.snr_weights(g["t"], ns.acp_host, 5.0), rtol=0)
    ns.prediction_type = "v_prediction"
    np.testing.assert_allclose(ns.snr_weights(t, 5.0), T.snr_weights(g["t"], ns.acp_host, 5.0, "v_prediction"), rtol=0)
    ts = ns.sample_timesteps(64, torch.Generator().manual_seed(0))
    assert ts.dtype == torch.int64 and ts.min() >= 0 and ts.max() < 1000


def test_grad_buckets_cover_the_buffer():
    b = sharding.grad_buckets(10, 4)
    assert [(r.start, r.stop) for r in b] == [(0, 4), (4, 8), (8, 10)]
    assert sharding.grad_buckets(0, 4) == []
    with pytest.raises(ValueError):
        sharding.grad_buckets(10, 0)
    # config 4: 618.8 M fp32 gradients in 64 Mi-element buckets
    assert len(sharding.grad_buckets(618_832_960, 64 << 20)) == 10


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        sharding.allreduce_flat_grads(g, bucket_elems=300)
        loss = sharding.gather_loss(torch.tensor([float(rank + 1)]))
        works = sharding.allreduce_flat_grads(torch.ones(10), bucket_elems=4, async_op=True)
        for w in works:
            w.wait()
        q.put((rank, g.clone(), loss, len(works)))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, g, loss, nworks in res:
        assert torch.equal(g, torch.arange(1000, dtype=torch.float32) * 3)   # SUM over ranks; the mean is the AdamW grad_scale
        assert loss == 1.5 and nworks == 3


def test_lr_schedules_match_the_reference_get_scheduler(golden_dir):
    from mirrorfusion_b200.train import LRSchedule
    g = np.load(os.path.join(golden_dir, "lr_schedules.npz"))
    for name in LRSchedule.NAMES:
        opt = B200AdamW(FlatParams({"w": (4,)}, "cpu", with_bf16=False), lr=1e-4)
        sch = LRSchedule(opt, name, num_warmup_steps=5, num_training_steps=40)
        lrs = []
        for _ in range(45):
            lrs.append(sch.get_last_lr()[0])
            sch.step()
        np.testing.assert_allclose(lrs, g[name], rtol=1e-12, atol=1e-18, err_msg=name)
    with pytest.raises(ValueError):
        LRSchedule(opt, "polynomial")


def test_brushnet_branch_pack_unpack_roundtrip():
    """Checkpoint export: flat packed layout -> the reference's state_dict naming / OIHW layout, bit-exact."""
    from mirrorfusion_b200.backward import brushnet_branch_shapes, pack_brushnet_branch, unpack_brushnet_branch
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200.synth import make_state_dict
    sd = make_state_dict(MICRO, "brushnet")
    flat = FlatParams(brushnet_branch_shapes(MICRO), "cpu", with_bf16=False)
    for k, v in pack_brushnet_branch(MICRO, sd).items():
        flat.p(k).copy_(v)
    back = unpack_brushnet_branch(MICRO, flat)
    assert len(back) >= 250
    for k, v in back.items():
        assert v.shape == sd[k].shape and torch.equal(v, sd[k].float()), k
    missing = [k for k in sd if k not in back]
    assert all(k.startswith(("conv_in_condition.", "time_embedding.")) or ".time_emb_proj." in k for k in missing)


def test_unpack_brushnet_round_trip_covers_every_trained_parameter():
    """the export must be the exact inverse of pack_brushnet over EVERY parameter BrushNetTrainer trains — branch,
    conv_in_condition (OIHW) and the timestep path (wcat / bcat split back per resnet) — and pass the strict census check."""
    import torch
    from mirrorfusion_b200 import checkpoint as CK
    from mirrorfusion_b200.backward import brushnet_shapes, pack_brushnet, unpack_brushnet
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    sd = make_state_dict(MICRO, "brushnet", seed=3)
    flat = FlatParams(brushnet_shapes(MICRO), "cpu", with_bf16=False)
    flat.load_state_dict(pack_brushnet(MICRO, sd))
    out = unpack_brushnet(MICRO, flat)
    CK.check_state_dict(out, MICRO, "brushnet")                 # names and shapes of the reference's BrushNetModel.state_dict()
    assert set(out) == set(sd)
    for k, v in sd.items():
        assert torch.equal(out[k], v), k
    # a "trained" buffer: every element changed -> every exported tensor changes (nothing is exported from stale copies)
    flat.param.add_(1.0)
    out2 = unpack_brushnet(MICRO, flat)
    assert all(not torch.equal(out2[k], sd[k]) for k in sd)
