"""Pins oracle/train_oracle.py (the CPU restatement of the fine-tune step's glue): against vectors produced by the reference's
OWN DDPMScheduler.add_noise / get_velocity / compute_snr and loss block (tests/golden/train_glue.npz, oracle/make_golden_train.py),
and against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW themselves (the calls E/train_brushnet_mirror.py:1460-1464 makes).
Also the host-only weight transform of the conv data gradient."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import train_oracle as T


def _gold(golden_dir):
    return np.load(os.path.join(golden_dir, "train_glue.npz"))


def test_noise_schedule_and_add_noise_match_reference(golden_dir):
    g = _gold(golden_dir)
    acp = T.alphas_cumprod()
    assert np.array_equal(acp, g["alphas_cumprod"])
    assert np.array_equal(T.add_noise(g["x0"], g["noise"], g["t"], acp), g["noisy"])
    assert np.array_equal(T.get_velocity(g["x0"], g["noise"], g["t"], acp), g["velocity"])
    np.testing.assert_allclose(T.compute_snr(g["t"], acp), g["snr"], rtol=1e-6)


def test_loss_and_gradient_match_reference_autograd(golden_dir):
    g = _gold(golden_dir)
    acp = T.alphas_cumprod()
    for name, w in (("plain", None), ("snr5", T.snr_weights(g["t"], acp, 5.0))):
        if w is not None:
            np.testing.assert_allclose(w, g["w_snr5"], rtol=1e-6)
        loss, per, grad = T.mse_loss(g["pred"], g["noise"], w)
        assert abs(loss - float(g[f"loss_{name}"])) < 2e-6 * abs(loss)
        np.testing.assert_allclose(grad, g[f"grad_{name}"], rtol=2e-5, atol=1e-9)


def test_clip_and_adamw_match_torch():
    gen = torch.Generator().manual_seed(5)
    shapes = [(7, 5), (33,), (4, 3, 3, 3)]
    params = [torch.nn.Parameter(torch.randn(s, generator=gen)) for s in shapes]
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    p = [q.detach().numpy().copy() for q in params]
    m = [np.zeros(s) for s in shapes]
    v = [np.zeros(s) for s in shapes]
    for step in range(1, 6):
        grads = [torch.randn(s, generator=gen) * (3.0 if step % 2 else 0.01) for s in shapes]
        for q, gr in zip(params, grads):
            q.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        coef, tot = T.clip_coef([gr.numpy() for gr in grads], 1.0)
        assert abs(tot - float(total)) < 1e-5 * tot
        for i in range(len(shapes)):
            p[i], m[i], v[i] = T.adamw_step(p[i], grads[i].numpy() * coef, m[i], v[i], step=step, lr=1e-3)
            np.testing.assert_allclose(p[i], params[i].detach().numpy(), rtol=3e-6, atol=1e-7)


def test_dgrad_weight_transform_is_the_data_gradient():
    # the packed dgrad weight, unpacked back to OIHW and applied as a FORWARD conv to dy, equals autograd's dx
    import mirrorfusion_b200.ops as ops
    gen = torch.Generator().manual_seed(9)
    for k in (3, 1):
        x = torch.randn(2, 16, 7, 9, generator=gen)
        w = torch.randn(24, 16, k, k, generator=gen)
        dy = torch.randn(2, 24, 7, 9, generator=gen)
        dx, dw, db = T.conv_grads(x.numpy(), w.numpy(), dy.numpy())
        with ops.precision("fp32"):
            wp = ops.pack_conv_dgrad_weight(w)                        # [Cin_fwd, k*k*Cout_fwd], K order (kh, kw, c)
        assert tuple(wp.shape) == (16, k * k * 24)
        w_back = wp.view(16, k, k, 24).permute(0, 3, 1, 2)
        got = F.conv2d(dy, w_back, padding=k // 2)
        np.testing.assert_allclose(got.numpy(), dx, rtol=1e-4, atol=1e-4)
        # and the wgrad K order the kernel writes: dw[co][(kh, kw, ci)]
        assert dw.transpose(0, 2, 3, 1).reshape(24, -1).shape == (24, k * k * 16)


def test_stride2_dgrad_weight_is_the_downsample_data_gradient():
    # evaluate the four phase convolutions of the packed weight with plain torch and interleave them: must equal autograd's dx
    import mirrorfusion_b200.ops as ops
    gen = torch.Generator().manual_seed(3)
    Cin, Cout, H, W = 6, 10, 8, 12
    x = torch.randn(2, Cin, H, W, generator=gen, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen, dtype=torch.float64)
    dy = torch.randn(2, Cout, H // 2, W // 2, generator=gen, dtype=torch.float64)
    F.conv2d(x, w, stride=2, padding=1).backward(dy)
    with ops.precision("fp32"):
        wp = ops.pack_conv_s2_dgrad_weight(w).double()                   # [4, Cin, 4*Cout]
    assert tuple(wp.shape) == (4, Cin, 4 * Cout)
    dx = torch.zeros_like(x)
    dyp = F.pad(dy, (1, 1, 1, 1))
    for py in range(2):
        for px in range(2):
            wk = wp[py * 2 + px].view(Cin, 2, 2, Cout).permute(0, 3, 1, 2)  # taps (ty, tx) at low-res offsets {-1,0} / {0,+1}
            y0 = 0 if py == 0 else 1
            x0 = 0 if px == 0 else 1
            win = dyp[:, :, y0:y0 + H // 2 + 1, x0:x0 + W // 2 + 1]
            dx[:, :, py::2, px::2] = F.conv2d(win, wk)
    np.testing.assert_allclose(dx.numpy(), x.grad.numpy(), rtol=1e-5, atol=1e-5)      # the packed weight is fp32


def _block_case(tag):
    cin, cout = (64, 64) if tag == "id" else (64, 128)
    return (cin, cout) + T.resnet_block_case(cin, cout)


@pytest.mark.parametrize("tag", ["id", "sc"])
def test_resnet_block_grads_oracle_vs_reference_autograd(golden_dir, tag):
    """oracle/train_oracle.resnet_block_grads against the reference's OWN ResnetBlock2D forward + autograd (fp32)."""
    g = np.load(os.path.join(golden_dir, "resnet_block_grad.npz"))
    cin, cout, sd, x, emb, d_out = _block_case(tag)
    r = T.resnet_block_grads(sd, "r", x, emb, d_out)
    for k in ("out", "dx", "d_rowbias"):
        np.testing.assert_allclose(r[k].numpy(), g[f"{tag}_{k}"], rtol=2e-4, atol=2e-5)
    for name in sd:
        short = name[2:]
        gr = r[name].reshape(-1)
        assert abs(float(gr.norm()) - float(g[f"{tag}_g_{short}_norm"])) < 1e-4 * float(g[f"{tag}_g_{short}_norm"]), name
        np.testing.assert_allclose(gr[torch.from_numpy(g[f"{tag}_g_{short}_idx"])].numpy(), g[f"{tag}_g_{short}_val"], rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("tag", ["id", "sc"])
def test_resnet_block_program_dataflow_on_the_cpu_stand_in(tag):
    """The launch program of mirrorfusion_b200/backward.py (forward + backward of one block) run on tests/torch_kernels.py —
    same call sequence, torch math — must reproduce the oracle's gradients: checks the PROGRAM (which buffer feeds which op,
    the residual path, accumulation into the flat gradient buffer), not the kernels."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import ResnetBlockTrainer, pack_resnet_state_dict, resnet_param_shapes, unpack_conv_grad
    from mirrorfusion_b200.train import FlatParams
    cin, cout, sd, x, emb, d_out = _block_case(tag)
    ref = T.resnet_block_grads(sd, "r", x, emb, d_out)
    B, _, H, W = x.shape
    flat = FlatParams(resnet_param_shapes("r", cin, cout), "cpu", with_bf16=False)
    for k, v in pack_resnet_state_dict("r", sd).items():
        flat.p(k).copy_(v)
    blk = ResnetBlockTrainer(flat, "r", B=B, H=H, W=W, Cin=cin, Cout=cout, precision="fp32", K=TK)
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(B, H * W, -1).contiguous()
    out = blk.forward(nhwc(x), ref["rowbias"].float())
    np.testing.assert_allclose(out.numpy(), nhwc(ref["out"]).numpy(), rtol=1e-4, atol=1e-4)
    for rep in range(2):                                   # gradients ACCUMULATE over micro-batches
        dx, drb = blk.backward(nhwc(d_out))
    np.testing.assert_allclose(dx.numpy(), nhwc(ref["dx"]).numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(drb.numpy(), ref["d_rowbias"].numpy(), rtol=1e-4, atol=1e-4)
    for name in flat.table:
        got = flat.g(name)
        want = ref[name]
        got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
        np.testing.assert_allclose(got.numpy(), 2 * want.numpy(), rtol=2e-4, atol=2e-4, err_msg=name)


def test_downsample_program_dataflow_on_the_cpu_stand_in():
    """DownsampleTrainer (forward stride-2 plan, weight gradient, data gradient through the up2x plan) on tests/torch_kernels.py
    against autograd of the oracle's Downsample2D (oracle/mf_oracle.py `downsample`)."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import DownsampleTrainer, unpack_conv_grad
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    gen = torch.Generator().manual_seed(21)
    B, C, H, W = 2, 16, 8, 12
    sd = {"d.conv.weight": (torch.randn(C, C, 3, 3, generator=gen) * 0.1).requires_grad_(True),
          "d.conv.bias": (torch.randn(C, generator=gen) * 0.1).requires_grad_(True)}
    x = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
    d_out = torch.randn(B, C, H // 2, W // 2, generator=gen)
    y = O.downsample(sd, "d", x)
    y.backward(d_out)
    flat = FlatParams({"d.conv.weight": (C, 9 * C), "d.conv.bias": (C,)}, "cpu", with_bf16=False)
    flat.p("d.conv.weight").copy_(sd["d.conv.weight"].detach().permute(0, 2, 3, 1).reshape(C, -1))
    flat.p("d.conv.bias").copy_(sd["d.conv.bias"].detach())
    blk = DownsampleTrainer(flat, "d", B=B, H=H, W=W, C=C, precision="fp32", K=TK)
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).reshape(B, -1, C).contiguous()
    np.testing.assert_allclose(blk.forward(nhwc(x)).numpy(), nhwc(y).numpy(), rtol=1e-5, atol=1e-5)
    dx = blk.backward(nhwc(d_out))
    np.testing.assert_allclose(dx.numpy(), nhwc(x.grad).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(unpack_conv_grad(flat.g("d.conv.weight"), 3).numpy(), sd["d.conv.weight"].grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(flat.g("d.conv.bias").numpy(), sd["d.conv.bias"].grad.numpy(), rtol=1e-4, atol=1e-5)


def test_brushnet_down_mid_program_dataflow_on_the_cpu_stand_in():
    """BrushNetDownMidTrainer (down blocks + mid block + 13 zero-conv taps; fan-out gradients summed in the zero-conv
    data-gradient epilogue) on tests/torch_kernels.py against autograd through the oracle's brushnet_forward (pinned to the
    reference's BrushNetModel by tests/golden/micro_step.npz): taps, d h0, every parameter gradient, and the row-bias gradients via
    d time_emb_proj.bias = sum_b d rowbias[b]."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import BrushNetDownMidTrainer, brushnet_down_mid_shapes, pack_brushnet_down_mid, unpack_conv_grad
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    cfg = MICRO
    B, H, W = 2, 16, 16          # 2x2 at the deepest level (at 1x1 a 1-channel GroupNorm group has zero variance)
    gen = torch.Generator().manual_seed(8)
    sd = {k: v.double().requires_grad_(True) for k, v in make_state_dict(cfg, "brushnet").items()}
    sample = torch.randn(B, cfg.in_channels, H, W, generator=gen, dtype=torch.float64)
    cond = torch.randn(B, cfg.conditioning_channels, H, W, generator=gen, dtype=torch.float64)
    t = torch.tensor([500, 20])
    down, mid, _ = O.brushnet_forward(sd, cfg, sample, t, cond)
    d_down = [torch.randn(d.shape, generator=gen, dtype=torch.float64) for d in down]
    d_mid = torch.randn(mid.shape, generator=gen, dtype=torch.float64)
    (sum((a * b).sum() for a, b in zip(down, d_down)) + (mid * d_mid).sum()).backward()

    with torch.no_grad():                                          # the trainer's inputs: conv_in_condition output and the row biases
        h0 = F.conv2d(torch.cat([sample, cond], 1), sd["conv_in_condition.weight"], sd["conv_in_condition.bias"], padding=1)
        emb = O.time_embed(sd, t, B, cfg.block_out_channels[0], torch.float64)
    flat = FlatParams(brushnet_down_mid_shapes(cfg), "cpu", with_bf16=False)
    for k, v in pack_brushnet_down_mid(cfg, {k: v.detach() for k, v in sd.items()}).items():
        flat.p(k).copy_(v)
    net = BrushNetDownMidTrainer(flat, cfg, B=B, H=H, W=W, precision="fp32", K=TK)
    rb = {p: F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"]).detach().float()
          for p in net.resnet_prefixes()}
    nhwc = lambda x: x.detach().permute(0, 2, 3, 1).reshape(B, -1, x.shape[1]).float().contiguous()
    taps, mid_tap = net.forward(nhwc(h0), rb)
    assert len(taps) == len(down) == 12
    for a, b in zip(taps + [mid_tap], down + [mid]):
        np.testing.assert_allclose(a.numpy(), nhwc(b).numpy(), rtol=2e-4, atol=2e-4)
    d_h0, d_rb = net.backward([nhwc(d) for d in d_down], nhwc(d_mid))
    # (+ 1e-2: gradients that are analytically zero — a bias in front of a 1-channel-per-group GroupNorm — compare as absolute)
    rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-2))
    # fp32 stand-in vs float64 autograd through ~20 layers with 1-channel GroupNorm groups over as few as 4 pixels: 1e-3 rel-L2
    # d h0 is the gradient at conv_in_condition's output: its sum over pixels is that conv's bias gradient
    assert rel(d_h0.sum((0, 1)), sd["conv_in_condition.bias"].grad) < 1e-3
    for name in flat.table:
        want = sd[name].grad
        got = flat.g(name)
        got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
        assert rel(got, want) < 1e-3, name
    for p, g in d_rb.items():
        assert rel(g.sum(0), sd[f"{p}.time_emb_proj.bias"].grad) < 1e-3, p


def test_upsample_program_dataflow_on_the_cpu_stand_in():
    """UpsampleTrainer (sub-pixel forward; weight gradient over the materialised x2 input; data gradient = stride-1 plan + 2x2
    sum-pool as a stride-2 plan with a 0/1 weight) on tests/torch_kernels.py against autograd of the oracle's Upsample2D."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import UpsampleTrainer, unpack_conv_grad
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    gen = torch.Generator().manual_seed(22)
    B, C, H, W = 2, 16, 6, 4
    sd = {"u.conv.weight": (torch.randn(C, C, 3, 3, generator=gen) * 0.1).requires_grad_(True),
          "u.conv.bias": (torch.randn(C, generator=gen) * 0.1).requires_grad_(True)}
    x = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
    d_out = torch.randn(B, C, 2 * H, 2 * W, generator=gen)
    y = O.upsample(sd, "u", x)
    y.backward(d_out)
    flat = FlatParams({"u.conv.weight": (C, 9 * C), "u.conv.bias": (C,)}, "cpu", with_bf16=False)
    flat.p("u.conv.weight").copy_(sd["u.conv.weight"].detach().permute(0, 2, 3, 1).reshape(C, -1))
    flat.p("u.conv.bias").copy_(sd["u.conv.bias"].detach())
    blk = UpsampleTrainer(flat, "u", B=B, H=H, W=W, C=C, precision="fp32", K=TK)
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).reshape(B, -1, C).contiguous()
    np.testing.assert_allclose(blk.forward(nhwc(x)).numpy(), nhwc(y).numpy(), rtol=1e-5, atol=1e-5)
    dx = blk.backward(nhwc(d_out))
    np.testing.assert_allclose(dx.numpy(), nhwc(x.grad).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(unpack_conv_grad(flat.g("u.conv.weight"), 3).numpy(), sd["u.conv.weight"].grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(flat.g("u.conv.bias").numpy(), sd["u.conv.bias"].grad.numpy(), rtol=1e-4, atol=1e-5)


def test_skip_resnet_program_dataflow_on_the_cpu_stand_in():
    """SkipResnetBlockTrainer (resnet over cat([x, skip]) of BrushNet's up blocks) on the stand-in against autograd of the oracle
    block evaluated on the concatenated input."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import (SkipResnetBlockTrainer, pack_skip_resnet_state_dict, skip_resnet_param_shapes,
                                            unpack_conv_grad)
    from mirrorfusion_b200.train import FlatParams
    C1, C2, cout = 64, 32, 64
    sd, x, emb, d_out = T.resnet_block_case(C1 + C2, cout, seed=78)
    ref = T.resnet_block_grads(sd, "r", x, emb, d_out)
    B, _, H, W = x.shape
    flat = FlatParams(skip_resnet_param_shapes("r", C1, C2, cout), "cpu", with_bf16=False)
    for k, v in pack_skip_resnet_state_dict("r", sd, C1).items():
        flat.p(k).copy_(v)
    blk = SkipResnetBlockTrainer(flat, "r", B=B, H=H, W=W, C1=C1, C2=C2, Cout=cout, precision="fp32", K=TK)
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(B, H * W, -1).contiguous()
    out = blk.forward(nhwc(x[:, :C1]), nhwc(x[:, C1:]), ref["rowbias"].float())
    np.testing.assert_allclose(out.numpy(), nhwc(ref["out"]).numpy(), rtol=1e-4, atol=1e-4)
    dx1, dx2, drb = blk.backward(nhwc(d_out))
    np.testing.assert_allclose(torch.cat([dx1, dx2], -1).numpy(), nhwc(ref["dx"]).numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(drb.numpy(), ref["d_rowbias"].numpy(), rtol=1e-4, atol=1e-4)
    got_sc = torch.cat([flat.g("r.conv_shortcut.weight.a"), flat.g("r.conv_shortcut.weight.b")], 1)
    np.testing.assert_allclose(got_sc.numpy(), ref["r.conv_shortcut.weight"][:, :, 0, 0].numpy(), rtol=2e-4, atol=2e-4)
    for name in flat.table:
        if "conv_shortcut.weight" in name:
            continue
        want = ref[name]
        got = flat.g(name)
        got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-4, atol=2e-4, err_msg=name)


def test_brushnet_branch_program_dataflow_on_the_cpu_stand_in():
    """The WHOLE BrushNet branch behind conv_in_condition (down, mid, up blocks, all 28 zero-conv taps) as one forward + backward
    launch program on the stand-in, against autograd through the oracle's brushnet_forward: every tap, d h0, every parameter
    gradient, every row-bias gradient.  Down-path hidden states have three consumers here (next block, zero-conv, skip)."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import BrushNetBranchTrainer, brushnet_branch_shapes, pack_brushnet_branch, unpack_conv_grad
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    cfg = MICRO
    B, H, W = 2, 16, 16
    gen = torch.Generator().manual_seed(9)
    sd = {k: v.double().requires_grad_(True) for k, v in make_state_dict(cfg, "brushnet").items()}
    sample = torch.randn(B, cfg.in_channels, H, W, generator=gen, dtype=torch.float64)
    cond = torch.randn(B, cfg.conditioning_channels, H, W, generator=gen, dtype=torch.float64)
    t = torch.tensor([700, 3])
    down, mid, up = O.brushnet_forward(sd, cfg, sample, t, cond)
    assert (len(down), len(up)) == (12, 15)
    rnd = lambda ts: [torch.randn(x.shape, generator=gen, dtype=torch.float64) for x in ts]
    d_down, d_up, d_mid = rnd(down), rnd(up), rnd([mid])[0]
    (sum((a * b).sum() for a, b in zip(down + up, d_down + d_up)) + (mid * d_mid).sum()).backward()
    with torch.no_grad():
        h0 = F.conv2d(torch.cat([sample, cond], 1), sd["conv_in_condition.weight"], sd["conv_in_condition.bias"], padding=1)
        emb = O.time_embed(sd, t, B, cfg.block_out_channels[0], torch.float64)
    shapes = brushnet_branch_shapes(cfg)
    flat = FlatParams(shapes, "cpu", with_bf16=False)
    for k, v in pack_brushnet_branch(cfg, {k: v.detach() for k, v in sd.items()}).items():
        flat.p(k).copy_(v)
    net = BrushNetBranchTrainer(flat, cfg, B=B, H=H, W=W, precision="fp32", K=TK)
    rb = {p: F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"]).detach().float()
          for p in net.resnet_prefixes()}
    assert len(rb) == 8 + 2 + 12
    nhwc = lambda x: x.detach().permute(0, 2, 3, 1).reshape(B, -1, x.shape[1]).float().contiguous()

    def rel(a, b):      # gradients that are analytically zero (a bias in front of a 1-channel-per-group GroupNorm): absolute
        a, b = a.double(), b.double()
        return float(a.norm()) if float(b.norm()) < 1e-9 else float((a - b).norm() / b.norm())

    td, tm, tu = net.forward(nhwc(h0), rb)
    for a, b in zip(td + [tm] + tu, down + [mid] + up):
        assert rel(a, nhwc(b)) < 1e-4
    d_h0, d_rb = net.backward([nhwc(d) for d in d_down], nhwc(d_mid), [nhwc(d) for d in d_up])
    assert rel(d_h0.sum((0, 1)), sd["conv_in_condition.bias"].grad) < 1e-3
    # every trainable tensor of the branch except conv_in_condition and the timestep path is in the flat buffer
    covered = {n[:-2] if n.endswith((".weight.a", ".weight.b")) else n for n in shapes}
    skipped = [k for k in sd if k not in covered]
    assert all(k.startswith(("conv_in_condition.", "time_embedding.")) or ".time_emb_proj." in k for k in skipped), skipped
    for name in shapes:
        if name.endswith(".weight.b"):
            continue
        if name.endswith(".weight.a"):
            want = sd[name[:-2]].grad[:, :, 0, 0]
            got = torch.cat([flat.g(name), flat.g(name[:-2] + ".b")], 1)
        else:
            want = sd[name].grad
            got = flat.g(name)
            got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
        assert rel(got, want) < 1e-3, name
    for p, g in d_rb.items():
        assert rel(g.sum(0), sd[f"{p}.time_emb_proj.bias"].grad) < 1e-3, p
    assert len(d_rb) == 22


def test_time_path_program_dataflow_on_the_cpu_stand_in():
    """TimePathTrainer (sinusoid -> TimestepEmbedding MLP -> all time_emb_proj as one GEMV; backward from the per-resnet row-bias
    gradients) on the stand-in against autograd of the oracle's time_embed + per-resnet projections."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import TimePathTrainer, pack_time_path, time_path_shapes
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    cfg, B = MICRO, 3
    gen = torch.Generator().manual_seed(4)
    sd = {k: v.double().requires_grad_(True) for k, v in make_state_dict(cfg, "brushnet").items()}
    prefixes = sorted({k[:-len(".time_emb_proj.weight")] for k in sd if k.endswith(".time_emb_proj.weight")})
    assert len(prefixes) == 22
    t = torch.tensor([999, 0, 421])
    emb = O.time_embed(sd, t, B, cfg.block_out_channels[0], torch.float64)
    rb = {p: F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"]) for p in prefixes}
    d_rb = {p: torch.randn(v.shape, generator=gen, dtype=torch.float64) for p, v in rb.items()}
    sum((rb[p] * d_rb[p]).sum() for p in prefixes).backward()
    flat = FlatParams(time_path_shapes(cfg, prefixes), "cpu", with_bf16=False)
    for k, v in pack_time_path(cfg, {k: v.detach() for k, v in sd.items()}, prefixes).items():
        flat.p(k).copy_(v)
    tp = TimePathTrainer(flat, cfg, prefixes, B=B, precision="fp32", K=TK)
    got = tp.forward(t)
    for p in prefixes:
        np.testing.assert_allclose(got[p].numpy(), rb[p].detach().numpy(), rtol=1e-4, atol=1e-5)
    tp.backward({p: g.float() for p, g in d_rb.items()})
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    for k in ("time_embedding.linear_1.weight", "time_embedding.linear_1.bias", "time_embedding.linear_2.weight", "time_embedding.linear_2.bias"):
        assert rel(flat.g(k), sd[k].grad) < 1e-4, k
    assert rel(flat.g("time_emb_proj.wcat"), torch.cat([sd[p + ".time_emb_proj.weight"].grad for p in prefixes], 0)) < 1e-4
    assert rel(flat.g("time_emb_proj.bcat"), torch.cat([sd[p + ".time_emb_proj.bias"].grad for p in prefixes], 0)) < 1e-4


def test_whole_brushnet_program_every_parameter_gradient_on_the_cpu_stand_in():
    """BrushNetTrainer: (sample, cond, timesteps) -> 28 taps and back, EVERY parameter of the BrushNetModel state_dict receives its
    gradient (conv_in_condition, timestep MLP, 22 time_emb_proj, 22 resnets, 3 + 3 samplers, 28 zero-convs) — compared with float64
    autograd through the oracle's brushnet_forward, no parameter excluded."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import BrushNetTrainer, brushnet_resnet_prefixes, brushnet_shapes, pack_brushnet, unpack_conv_grad
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    cfg = MICRO
    B, H, W = 2, 16, 16
    gen = torch.Generator().manual_seed(10)
    sd = {k: v.double().requires_grad_(True) for k, v in make_state_dict(cfg, "brushnet").items()}
    sample = torch.randn(B, cfg.in_channels, H, W, generator=gen, dtype=torch.float64)
    cond = torch.randn(B, cfg.conditioning_channels, H, W, generator=gen, dtype=torch.float64)
    t = torch.tensor([850, 12])
    down, mid, up = O.brushnet_forward(sd, cfg, sample, t, cond)
    rnd = lambda ts: [torch.randn(x.shape, generator=gen, dtype=torch.float64) for x in ts]
    d_down, d_up, d_mid = rnd(down), rnd(up), rnd([mid])[0]
    (sum((a * b).sum() for a, b in zip(down + up, d_down + d_up)) + (mid * d_mid).sum()).backward()
    shapes = brushnet_shapes(cfg)
    flat = FlatParams(shapes, "cpu", with_bf16=False)
    for k, v in pack_brushnet(cfg, {k: v.detach() for k, v in sd.items()}).items():
        flat.p(k).copy_(v)
    assert flat.numel >= sum(v.numel() for v in sd.values())           # every parameter lives in the flat buffer (+ alignment padding)
    net = BrushNetTrainer(flat, cfg, B=B, H=H, W=W, precision="fp32", K=TK)
    nhwc = lambda x: x.detach().permute(0, 2, 3, 1).reshape(B, -1, x.shape[1]).float().contiguous()

    def rel(a, b):
        a, b = a.double(), b.double()
        return float(a.norm()) if float(b.norm()) < 1e-9 else float((a - b).norm() / b.norm())

    td, tm, tu = net.forward(sample.float(), cond.float(), t)
    for a, b in zip(td + [tm] + tu, down + [mid] + up):
        assert rel(a, nhwc(b)) < 1e-4
    net.backward([nhwc(d) for d in d_down], nhwc(d_mid), [nhwc(d) for d in d_up])
    prefixes = brushnet_resnet_prefixes(cfg)
    checked = set()
    for name in shapes:
        if name.endswith(".weight.b"):
            continue
        if name == "time_emb_proj.wcat":
            want = torch.cat([sd[p + ".time_emb_proj.weight"].grad for p in prefixes], 0)
            got = flat.g(name)
            checked |= {p + ".time_emb_proj.weight" for p in prefixes}
        elif name == "time_emb_proj.bcat":
            want = torch.cat([sd[p + ".time_emb_proj.bias"].grad for p in prefixes], 0)
            got = flat.g(name)
            checked |= {p + ".time_emb_proj.bias" for p in prefixes}
        elif name.endswith(".weight.a"):
            want = sd[name[:-2]].grad[:, :, 0, 0]
            got = torch.cat([flat.g(name), flat.g(name[:-2] + ".b")], 1)
            checked.add(name[:-2])
        else:
            want = sd[name].grad
            got = flat.g(name)
            got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
            checked.add(name)
        assert rel(got, want) < 1e-3, name
    assert checked == set(sd.keys())                                   # no parameter of the reference state_dict left out


def test_next_backward_algorithms_match_autograd():
    """The kernel-shaped restatements of attention / LayerNorm / GEGLU backward (the frozen UNet's dgrad chain, next to be built)
    against torch autograd of the functions the reference calls."""
    gen = torch.Generator().manual_seed(31)
    Tq, S, d = 37, 77, 40                                     # ragged: 37 queries against the 77-token CLIP context, head dim 40
    q, k, v, do = (torch.randn(n, d, generator=gen, dtype=torch.float64) for n in (Tq, S, S, Tq))
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    F.scaled_dot_product_attention(q[None, None], k[None, None], v[None, None])[0, 0].backward(do)
    dq, dk, dv, L, D = T.attention_backward_two_pass(q.detach().numpy(), k.detach().numpy(), v.detach().numpy(), do.numpy())
    for got, want in ((dq, q.grad), (dk, k.grad), (dv, v.grad)):
        np.testing.assert_allclose(got, want.numpy(), rtol=1e-9, atol=1e-11)

    x = torch.randn(5, 9, 320, generator=gen, dtype=torch.float64, requires_grad=True)
    gamma, beta = torch.randn(320, generator=gen, dtype=torch.float64), torch.randn(320, generator=gen, dtype=torch.float64)
    dy = torch.randn(5, 9, 320, generator=gen, dtype=torch.float64)
    F.layer_norm(x, (320,), gamma, beta, 1e-5).backward(dy)
    np.testing.assert_allclose(T.layernorm_backward_dx(x.detach().numpy(), gamma.numpy(), dy.numpy()), x.grad.numpy(), rtol=1e-9, atol=1e-11)

    proj = torch.randn(4, 6, 128, generator=gen, dtype=torch.float64, requires_grad=True)
    dout = torch.randn(4, 6, 64, generator=gen, dtype=torch.float64)
    h, gate = proj.chunk(2, -1)
    (h * F.gelu(gate)).backward(dout)
    np.testing.assert_allclose(T.geglu_backward(proj.detach().numpy(), dout.numpy()), proj.grad.numpy(), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("tag", ["id", "sc"])
def test_frozen_resnet_block_carries_only_data_gradients(tag):
    """frozen=True (a block of the frozen UNet): same d x / d rowbias, no parameter gradient touched."""
    import torch_kernels as TK
    from mirrorfusion_b200.backward import ResnetBlockTrainer, pack_resnet_state_dict, resnet_param_shapes
    from mirrorfusion_b200.train import FlatParams
    cin, cout, sd, x, emb, d_out = _block_case(tag)
    ref = T.resnet_block_grads(sd, "r", x, emb, d_out)
    B, _, H, W = x.shape
    flat = FlatParams(resnet_param_shapes("r", cin, cout), "cpu", with_bf16=False)
    for k, v in pack_resnet_state_dict("r", sd).items():
        flat.p(k).copy_(v)
    blk = ResnetBlockTrainer(flat, "r", B=B, H=H, W=W, Cin=cin, Cout=cout, precision="fp32", K=TK, frozen=True)
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(B, H * W, -1).contiguous()
    blk.forward(nhwc(x), ref["rowbias"].float())
    dx, drb = blk.backward(nhwc(d_out))
    np.testing.assert_allclose(dx.numpy(), nhwc(ref["dx"]).numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(drb.numpy(), ref["d_rowbias"].numpy(), rtol=1e-4, atol=1e-4)
    assert float(flat.grad.abs().max()) == 0.0
