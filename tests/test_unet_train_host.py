"""DATAFLOW of the frozen-UNet fine-tune program (mirrorfusion_b200/unet_train.py) on the CPU stand-in kernels
(tests/torch_kernels.py), against float64 autograd through the oracle's `unet_forward` — itself pinned to the reference's
UNet2DConditionModel with the BrushNet tap sites (tests/golden/*step*.npz).  What is checked is the PROGRAM: which tensors every
backward op reads, how a hidden state's two gradients (next block + skip) are summed inside the consumer's last kernel, where the 28
tap gradients live, the two-source GroupNorm of the skip concat, the un-fused GEGLU, cross attention without context gradients.
The kernels themselves are checked on the GPU (tests/test_gpu_train_bf16.py, tests/test_gpu_zz_unet_train.py)."""
import pytest
import torch

from mirrorfusion_b200.config import MICRO, tap_channels
from mirrorfusion_b200.synth import make_inputs, make_state_dict
from mirrorfusion_b200.unet_train import FrozenUNetTrainer
from oracle import mf_oracle as O

import torch_kernels as TK


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def nhwc(t):
    return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1, t.shape[1]).contiguous()


def _tap_shapes(cfg, B, H, W):
    dch, mch, uch = tap_channels(cfg)
    n = len(cfg.block_out_channels)
    hw, out = (H, W), []
    out.append((dch[0], hw))
    k = 1
    for i in range(n):
        for _ in range(cfg.layers_per_block):
            out.append((dch[k], hw)); k += 1
        if i != n - 1:
            hw = (hw[0] // 2, hw[1] // 2)
            out.append((dch[k], hw)); k += 1
    out.append((mch, hw))
    k = 0
    for i in range(n):
        for _ in range(cfg.layers_per_block + 1):
            out.append((uch[k], hw)); k += 1
        if i != n - 1:
            hw = (hw[0] * 2, hw[1] * 2)
            out.append((uch[k], hw)); k += 1
    return out


@pytest.mark.timeout(600)
def test_frozen_unet_program_tap_gradients_on_the_cpu_stand_in():
    cfg, B, H, W = MICRO, 2, 16, 16
    torch.manual_seed(0)
    sd = make_state_dict(cfg, "unet", seed=2)
    inp = make_inputs(cfg, B, seed=9, height=H, width=W, cfg_duplicate=False)
    g = torch.Generator().manual_seed(4)
    shapes = _tap_shapes(cfg, B, H, W)
    assert len(shapes) == 28
    taps_nchw = [0.3 * torch.randn(B, c, h, w, generator=g) for c, (h, w) in shapes]
    taps = [nhwc(t) for t in taps_nchw]
    sample = inp["latents"]
    tsteps = torch.tensor([37.0, 811.0])
    ehs = inp["prompt_embeds"]
    net = FrozenUNetTrainer(cfg, sd, taps, B=B, H=H, W=W, device="cpu", precision="fp32", K=TK)
    pred = net.forward(sample, tsteps, ehs)
    # float64 reference with autograd through the taps
    sd64 = {k: v.double() for k, v in sd.items()}
    t64 = [t.double().requires_grad_(True) for t in taps_nchw]
    nd = len(tap_channels(cfg)[0])
    ref = O.unet_forward(sd64, cfg, sample.double(), tsteps.double(), ehs.double(), t64[:nd], t64[nd], t64[nd + 1:])
    assert rel(pred, ref) < 1e-5
    d_pred = torch.randn(ref.shape, generator=g)
    ref.backward(d_pred.double())
    dd, dm, du = net.backward(d_pred)
    got = list(dd) + [dm] + list(du)
    assert len(got) == 28
    for k, (a, t) in enumerate(zip(got, t64)):
        assert a.shape == taps[k].shape
        assert rel(a, nhwc(t.grad)) < 1e-3, k      # fp32 stand-in vs float64 through ~60 layers (2x2 maps at the bottom)
    # a second step with other inputs reuses every buffer
    pred2 = net.forward(sample * 0.5, tsteps + 3, ehs * 0.9)
    ref2 = O.unet_forward(sd64, cfg, (sample * 0.5).double(), (tsteps + 3).double(), (ehs * 0.9).double(), [t.detach() for t in t64[:nd]],
                          t64[nd].detach(), [t.detach() for t in t64[nd + 1:]])
    assert rel(pred2, ref2) < 1e-5
    assert net.flops_fwd > 0 and net.flops_bwd > 0


@pytest.mark.timeout(900)
def test_fine_tune_dataflow_brushnet_through_frozen_unet_on_the_cpu_stand_in():
    """The two programs chained as FineTuneStep chains them (tap buffers shared forward, tap-gradient buffers bound backward):
    loss -> frozen UNet data gradients -> 28 taps -> EVERY BrushNet parameter gradient, against float64 autograd of the oracle
    (E/train_brushnet_mirror.py:836-888,1433-1459)."""
    import torch.nn.functional as F
    from mirrorfusion_b200.backward import (BrushNetTrainer, brushnet_resnet_prefixes, brushnet_shapes, pack_brushnet, unpack_conv_grad)
    from mirrorfusion_b200.train import FlatParams
    cfg, B, H, W = MICRO, 2, 16, 16
    usd, bsd = make_state_dict(cfg, "unet", seed=2), make_state_dict(cfg, "brushnet", seed=2)
    inp = make_inputs(cfg, B, seed=9, height=H, width=W, cfg_duplicate=False)
    g = torch.Generator().manual_seed(4)
    noisy, target = torch.randn(B, 4, H, W, generator=g), torch.randn(B, 4, H, W, generator=g)
    tsteps = torch.tensor([37, 811])
    cond, ehs = inp["conditioning_latents"], inp["prompt_embeds"]
    flat = FlatParams(brushnet_shapes(cfg), "cpu", with_bf16=False)
    flat.load_state_dict(pack_brushnet(cfg, bsd))
    bn = BrushNetTrainer(flat, cfg, B=B, H=H, W=W, precision="fp32", K=TK)
    br = bn.branch
    taps = [z.tap for z in br.taps] + [br.mid_tap.tap] + [z.tap for z in br.up_taps]
    unet = FrozenUNetTrainer(cfg, usd, taps, B=B, H=H, W=W, device="cpu", precision="fp32", K=TK)
    bn.bind_tap_gradients(unet.d_taps)
    bn.forward(noisy, cond, tsteps)
    pred = unet.forward(noisy, tsteps, ehs)
    d_pred = 2.0 * (pred - target) / pred.numel()
    dd, dm, du = unet.backward(d_pred)
    bn.backward(dd, dm, du)
    bs = {k: v.double().requires_grad_(True) for k, v in bsd.items()}
    us = {k: v.double() for k, v in usd.items()}
    down, mid, up = O.brushnet_forward(bs, cfg, noisy.double(), tsteps, cond.double())
    ref = O.unet_forward(us, cfg, noisy.double(), tsteps.double(), ehs.double(), down, mid, up)
    assert rel(pred, ref) < 1e-5
    F.mse_loss(ref, target.double()).backward()
    worst = 0.0
    for name in flat.table:
        got = flat.g(name)
        if name.endswith(".conv_shortcut.weight.a") or name.endswith(".conv_shortcut.weight.b"):
            full = bs[name[:-2]].grad[:, :, 0, 0]
            c1 = flat.shapes[name[:-2] + ".a"][1]
            want = full[:, :c1] if name.endswith(".a") else full[:, c1:]
        elif name in ("time_emb_proj.wcat", "time_emb_proj.bcat"):
            suffix = ".time_emb_proj.weight" if name.endswith("wcat") else ".time_emb_proj.bias"
            want = torch.cat([bs[p + suffix].grad for p in brushnet_resnet_prefixes(cfg)], 0)
        else:
            want = bs[name].grad
            if want.dim() == 4:
                got = unpack_conv_grad(got, want.shape[-1]) if want.shape[-1] == 3 else got.reshape(want.shape)
        if float(want.norm()) < 1e-9:      # MICRO has one channel per GroupNorm group: a conv bias in front of it has zero gradient
            assert float(got.norm()) < 1e-5, name
            continue
        e = rel(got.reshape(want.shape), want)
        worst = max(worst, e)
        assert e < 2e-3, (name, e)
    assert worst > 0
