"""Net-level backward program on the GPU (runs last): BrushNet's down path + mid block + 13 zero-conv taps
(mirrorfusion_b200/backward.BrushNetDownMidTrainer) in fp32 parity mode against float64 autograd through the oracle's
brushnet_forward.  The same program is checked on the CPU stand-in in tests/test_oracle_train.py."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_brushnet_down_mid_forward_backward_fp32_vs_autograd():
    from mirrorfusion_b200 import ops
    from mirrorfusion_b200.backward import BrushNetDownMidTrainer, brushnet_down_mid_shapes, pack_brushnet_down_mid, unpack_conv_grad
    from mirrorfusion_b200.config import TINY
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    ops.lib()
    cfg = TINY
    B, H, W = 2, 16, 16
    gen = torch.Generator().manual_seed(8)
    sd = {k: v.double().requires_grad_(True) for k, v in make_state_dict(cfg, "brushnet").items()}
    sample = torch.randn(B, cfg.in_channels, H, W, generator=gen, dtype=torch.float64)
    cond = torch.randn(B, cfg.conditioning_channels, H, W, generator=gen, dtype=torch.float64)
    t = torch.tensor([500, 20])
    down, mid, _ = O.brushnet_forward(sd, cfg, sample, t, cond)
    d_down = [torch.randn(d.shape, generator=gen, dtype=torch.float64) for d in down]
    d_mid = torch.randn(mid.shape, generator=gen, dtype=torch.float64)
    (sum((a * b).sum() for a, b in zip(down, d_down)) + (mid * d_mid).sum()).backward()
    with torch.no_grad():
        h0 = F.conv2d(torch.cat([sample, cond], 1), sd["conv_in_condition.weight"], sd["conv_in_condition.bias"], padding=1)
        emb = O.time_embed(sd, t, B, cfg.block_out_channels[0], torch.float64)
    flat = FlatParams(brushnet_down_mid_shapes(cfg), "cuda")
    flat.load_state_dict(pack_brushnet_down_mid(cfg, {k: v.detach() for k, v in sd.items()}))
    net = BrushNetDownMidTrainer(flat, cfg, B=B, H=H, W=W, precision="fp32")
    rb = {p: F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"]).detach().float().cuda()
          for p in net.resnet_prefixes()}
    nhwc = lambda x: x.detach().permute(0, 2, 3, 1).reshape(B, -1, x.shape[1]).float().contiguous()
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-2))
    taps, mid_tap = net.forward(nhwc(h0).cuda(), rb)
    for a, b in zip(taps + [mid_tap], down + [mid]):
        assert rel(a, nhwc(b)) < 1e-4
    d_h0, d_rb = net.backward([nhwc(d).cuda() for d in d_down], nhwc(d_mid).cuda())
    # fp32 kernels vs float64 autograd through 22 layers (8 / 16 channels per GroupNorm group, maps down to 2x2): 1e-3 rel-L2
    assert rel(d_h0.sum((0, 1)), sd["conv_in_condition.bias"].grad) < 1e-3
    for name in flat.table:
        want = sd[name].grad
        got = flat.g(name)
        got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
        assert rel(got, want) < 1e-3, name
    for p, g in d_rb.items():
        assert rel(g.sum(0), sd[f"{p}.time_emb_proj.bias"].grad) < 1e-3, p


@pytest.mark.timeout(180)
def test_brushnet_whole_branch_forward_backward_fp32_vs_autograd():
    """All 28 taps: BrushNetBranchTrainer in fp32 parity mode against float64 autograd through the oracle's brushnet_forward."""
    from mirrorfusion_b200 import ops
    from mirrorfusion_b200.backward import BrushNetBranchTrainer, brushnet_branch_shapes, pack_brushnet_branch, unpack_conv_grad
    from mirrorfusion_b200.config import TINY
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    ops.lib()
    cfg = TINY
    B, H, W = 2, 16, 16
    gen = torch.Generator().manual_seed(9)
    sd = {k: v.double().requires_grad_(True) for k, v in make_state_dict(cfg, "brushnet").items()}
    sample = torch.randn(B, cfg.in_channels, H, W, generator=gen, dtype=torch.float64)
    cond = torch.randn(B, cfg.conditioning_channels, H, W, generator=gen, dtype=torch.float64)
    t = torch.tensor([700, 3])
    down, mid, up = O.brushnet_forward(sd, cfg, sample, t, cond)
    rnd = lambda ts: [torch.randn(x.shape, generator=gen, dtype=torch.float64) for x in ts]
    d_down, d_up, d_mid = rnd(down), rnd(up), rnd([mid])[0]
    (sum((a * b).sum() for a, b in zip(down + up, d_down + d_up)) + (mid * d_mid).sum()).backward()
    with torch.no_grad():
        h0 = F.conv2d(torch.cat([sample, cond], 1), sd["conv_in_condition.weight"], sd["conv_in_condition.bias"], padding=1)
        emb = O.time_embed(sd, t, B, cfg.block_out_channels[0], torch.float64)
    shapes = brushnet_branch_shapes(cfg)
    flat = FlatParams(shapes, "cuda")
    flat.load_state_dict(pack_brushnet_branch(cfg, {k: v.detach() for k, v in sd.items()}))
    net = BrushNetBranchTrainer(flat, cfg, B=B, H=H, W=W, precision="fp32")
    rb = {p: F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"]).detach().float().cuda()
          for p in net.resnet_prefixes()}
    nhwc = lambda x: x.detach().permute(0, 2, 3, 1).reshape(B, -1, x.shape[1]).float().contiguous()

    def rel(a, b):
        a, b = a.double().cpu(), b.double()
        return float(a.norm()) if float(b.norm()) < 1e-9 else float((a - b).norm() / b.norm())

    td, tm, tu = net.forward(nhwc(h0).cuda(), rb)
    for a, b in zip(td + [tm] + tu, down + [mid] + up):
        assert rel(a, nhwc(b)) < 1e-4
    d_h0, d_rb = net.backward([nhwc(d).cuda() for d in d_down], nhwc(d_mid).cuda(), [nhwc(d).cuda() for d in d_up])
    assert rel(d_h0.sum((0, 1)), sd["conv_in_condition.bias"].grad) < 1e-3
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


@pytest.mark.timeout(180)
def test_whole_brushnet_every_parameter_gradient_fp32_vs_autograd():
    """BrushNetTrainer in fp32 parity mode: (sample, cond, timesteps) -> 28 taps and back; every parameter of the BrushNetModel
    state_dict against float64 autograd through the oracle's brushnet_forward."""
    from mirrorfusion_b200 import ops
    from mirrorfusion_b200.backward import BrushNetTrainer, brushnet_resnet_prefixes, brushnet_shapes, pack_brushnet, unpack_conv_grad
    from mirrorfusion_b200.config import TINY
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    ops.lib()
    cfg = TINY
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
    flat = FlatParams(shapes, "cuda")
    flat.load_state_dict(pack_brushnet(cfg, {k: v.detach() for k, v in sd.items()}))
    net = BrushNetTrainer(flat, cfg, B=B, H=H, W=W, precision="fp32")
    nhwc = lambda x: x.detach().permute(0, 2, 3, 1).reshape(B, -1, x.shape[1]).float().contiguous()

    def rel(a, b):
        a, b = a.double().cpu(), b.double()
        return float(a.norm()) if float(b.norm()) < 1e-9 else float((a - b).norm() / b.norm())

    td, tm, tu = net.forward(sample.float().cuda(), cond.float().cuda(), t.cuda())
    for a, b in zip(td + [tm] + tu, down + [mid] + up):
        assert rel(a, nhwc(b)) < 1e-4
    net.backward([nhwc(d).cuda() for d in d_down], nhwc(d_mid).cuda(), [nhwc(d).cuda() for d in d_up])
    prefixes = brushnet_resnet_prefixes(cfg)
    for name in shapes:
        if name.endswith(".weight.b"):
            continue
        if name == "time_emb_proj.wcat":
            want, got = torch.cat([sd[p + ".time_emb_proj.weight"].grad for p in prefixes], 0), flat.g(name)
        elif name == "time_emb_proj.bcat":
            want, got = torch.cat([sd[p + ".time_emb_proj.bias"].grad for p in prefixes], 0), flat.g(name)
        elif name.endswith(".weight.a"):
            want, got = sd[name[:-2]].grad[:, :, 0, 0], torch.cat([flat.g(name), flat.g(name[:-2] + ".b")], 1)
        else:
            want, got = sd[name].grad, flat.g(name)
            got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
        assert rel(got, want) < 1e-3, name


@pytest.mark.timeout(120)
def test_layernorm_and_geglu_backward_f32():
    import numpy as np
    from mirrorfusion_b200 import ops
    from oracle import train_oracle as T
    ops.lib()
    g = torch.Generator().manual_seed(1)
    x, dy = torch.randn(37, 320, generator=g), torch.randn(37, 320, generator=g)
    gamma = 1 + 0.2 * torch.randn(320, generator=g)
    dx = torch.full_like(x, float("nan")).cuda()
    ops.layernorm_bwd_f32(x.cuda(), dy.cuda(), gamma.cuda(), dx)
    ref = T.layernorm_backward_dx(x.numpy(), gamma.numpy(), dy.numpy())
    assert np.linalg.norm(dx.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-5
    proj, dout = torch.randn(50, 256, generator=g), torch.randn(50, 128, generator=g)
    out, dproj = torch.full((50, 128), float("nan")).cuda(), torch.full((50, 256), float("nan")).cuda()
    ops.geglu_f32(proj.cuda(), out=out, d_out=dout.cuda(), d_proj=dproj)
    h, gate = proj.chunk(2, -1)
    assert torch.allclose(out.cpu(), h * F.gelu(gate), rtol=1e-5, atol=1e-6)
    ref = T.geglu_backward(proj.numpy(), dout.numpy())
    assert np.linalg.norm(dproj.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-5


@pytest.mark.timeout(120)
def test_attention_backward_f32():
    import numpy as np
    from mirrorfusion_b200 import ops
    from oracle import train_oracle as T
    ops.lib()
    g = torch.Generator().manual_seed(2)
    B, heads, d, Tq, Tk = 2, 2, 40, 37, 77
    q, do = torch.randn(B, Tq, heads * d, generator=g), torch.randn(B, Tq, heads * d, generator=g)
    k, v = torch.randn(B, Tk, heads * d, generator=g), torch.randn(B, Tk, heads * d, generator=g)
    dq, dk, dv = (torch.full_like(t, float("nan")).cuda() for t in (q, k, v))
    ws = torch.zeros(2 * B * heads * Tq).cuda()
    ops.attention_bwd_f32(q.cuda(), k.cuda(), v.cuda(), do.cuda(), dq, dk, dv, ws, B=B, heads=heads, head_dim=d, Tq=Tq, Tk=Tk)
    for b in range(B):
        for h in range(heads):
            sl = slice(h * d, (h + 1) * d)
            rq, rk, rv, _, _ = T.attention_backward_two_pass(q[b, :, sl].numpy(), k[b, :, sl].numpy(), v[b, :, sl].numpy(), do[b, :, sl].numpy())
            for got, want in ((dq, rq), (dk, rk), (dv, rv)):
                got = got[b, :, sl].cpu().numpy()
                assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5
