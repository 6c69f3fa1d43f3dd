"""bf16 backward kernels of the fine-tune step (BASELINE config 4) through the C ABI, against float32 / float64 torch autograd of
the same op on the same (bf16-rounded) inputs: tcgen05 flash-attention backward (csrc/attn_bwd.cu), GroupNorm(+SiLU), LayerNorm,
GEGLU and conv_out data gradients (csrc/train_bf16.cu).  Tolerances are bf16 storage of the results (2^-9 per element)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from mirrorfusion_b200 import ops as o
    o.lib()
    return o


def randn(*shape, seed=0, dtype=bf16, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("B,heads,d,Tq,Tk,cross", [
    (2, 8, 40, 1024, 1024, False), (1, 8, 40, 4096, 4096, False), (2, 8, 80, 1024, 1024, False), (2, 8, 160, 256, 256, False),
    (2, 8, 160, 64, 64, False), (2, 2, 32, 256, 256, False), (2, 2, 64, 128, 128, False), (1, 8, 40, 200, 333, False),
    (1, 4, 80, 130, 70, False), (2, 8, 40, 1024, 77, True), (2, 8, 80, 256, 77, True), (2, 8, 160, 64, 77, True)])
def test_attention_backward_vs_autograd(ops, B, heads, d, Tq, Tk, cross):
    """q / k / v read out of a fused [B, T, 3C] projection buffer (self attention) exactly as the UNet's transformer blocks keep
    them; gradients written into a [B, T, 3C] buffer the same way.  Cross attention: separate tensors, no dk / dv."""
    C = heads * d
    if cross:
        q, k, v = randn(B, Tq, C, seed=1), randn(B, Tk, C, seed=2), randn(B, Tk, C, seed=3)
        ld = dict(ldq=C, ldk=C, ldv=C)
    else:
        assert Tq == Tk or True
        qkv = randn(B, max(Tq, Tk), 3 * C, seed=1)
        q, k, v = qkv[:, :Tq, :C], qkv[:, :Tk, C:2 * C], qkv[:, :Tk, 2 * C:]
        if Tq != Tk:      # views with different lengths: use contiguous copies (the leading dimension stays 3C-like only when equal)
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
            ld = dict(ldq=C, ldk=C, ldv=C)
        else:
            ld = dict(ldq=3 * C, ldk=3 * C, ldv=3 * C)
    d_o = randn(B, Tq, C, seed=4)
    o = torch.full((B, Tq, C), float("nan"), device="cuda", dtype=bf16)
    lse = torch.zeros(B * heads * Tq, device="cuda")
    ops.attention_lse(q, k, v, o, lse, B=B, heads=heads, head_dim=d, Tq=Tq, Tk=Tk, **ld)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(B, Tq, heads, d).transpose(1, 2)
    kh = kf.view(B, Tk, heads, d).transpose(1, 2)
    vh = vf.view(B, Tk, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * d ** -0.5
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Tq, C)
    assert rel(o.float(), ref) < 6e-3
    want_lse = (torch.logsumexp(s, -1) * 1.4426950408889634).reshape(-1)            # log2 domain
    assert (lse - want_lse).abs().max().item() < 2e-2
    ref.backward(d_o.float())
    dvec = torch.zeros(B * heads * Tq, device="cuda")
    if cross or Tq != Tk:
        dq = torch.full((B, Tq, C), float("nan"), device="cuda", dtype=bf16)
        dk = None if cross else torch.full((B, Tk, C), float("nan"), device="cuda", dtype=bf16)
        dv = None if cross else torch.full((B, Tk, C), float("nan"), device="cuda", dtype=bf16)
        ops.attention_bwd(q, k, v, o, d_o, lse, dvec, dq, dk, dv, B=B, heads=heads, head_dim=d, Tq=Tq, Tk=Tk, **ld)
    else:
        dqkv = torch.full((B, Tq, 3 * C), float("nan"), device="cuda", dtype=bf16)
        dq, dk, dv = dqkv[:, :, :C], dqkv[:, :, C:2 * C], dqkv[:, :, 2 * C:]
        ops.attention_bwd(q, k, v, o, d_o, lse, dvec, dq, dk, dv, B=B, heads=heads, head_dim=d, Tq=Tq, Tk=Tk, lddq=3 * C, lddk=3 * C,
                          lddv=3 * C, **ld)
    want_D = (d_o.float() * ref.detach()).view(B, Tq, heads, d).sum(-1).transpose(1, 2).reshape(-1)
    assert rel(dvec, want_D) < 2e-2            # O is bf16-rounded on our side
    assert torch.isfinite(dq.float()).all()
    assert rel(dq.float(), qf.grad) < 1.5e-2
    if not cross:
        assert torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()
        assert rel(dk.float(), kf.grad) < 1.5e-2
        assert rel(dv.float(), vf.grad) < 1.5e-2


@pytest.mark.parametrize("B,HW,C1,C2,groups,silu,eps", [
    (2, 64 * 64, 320, 0, 32, True, 1e-5), (2, 32 * 32, 640, 320, 32, True, 1e-5), (3, 16 * 16, 1280, 1280, 32, True, 1e-5),
    (2, 8 * 8, 1280, 0, 32, True, 1e-5), (2, 32 * 32, 640, 0, 32, False, 1e-6), (2, 15 * 9, 64, 128, 8, True, 1e-5),
    (4, 4 * 4, 128, 0, 8, True, 1e-5)])
def test_groupnorm_backward_vs_autograd(ops, B, HW, C1, C2, groups, silu, eps):
    C = C1 + C2
    x1 = randn(B, HW, C1, seed=1, scale=2.0) + 0.5
    x2 = None if C2 == 0 else randn(B, HW, C2, seed=2, scale=0.7)
    dy, dres, dres2 = randn(B, HW, C, seed=3), randn(B, HW, C, seed=4), randn(B, HW, C, seed=5)
    gamma = 1 + 0.2 * randn(C, seed=6, dtype=torch.float32)
    beta = 0.1 * randn(C, seed=7, dtype=torch.float32)
    xcat = (x1 if x2 is None else torch.cat([x1, x2], -1)).float().requires_grad_(True)
    g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xcat.transpose(1, 2), groups, g_, b_, eps).transpose(1, 2)
    if silu:
        y = F.silu(y)
    y.backward(dy.float())
    want_dx = xcat.grad + dres.float() + dres2.float()
    dx1 = torch.full_like(x1, float("nan"))
    dx2 = None if x2 is None else torch.full_like(x2, float("nan"))
    dg = torch.full((C,), 3.0, device="cuda")
    db = torch.full((C,), -2.0, device="cuda")
    for _ in range(2):      # twice: the ticket counters must re-arm
        ops.groupnorm_bwd(x1, x2, dy, gamma, beta, dx1, dx2, None, B=B, HW=HW, groups=groups, eps=eps, silu=silu, dgamma=dg, dbeta=db,
                          accumulate=True, dres=dres, dres2=dres2)
    got = dx1 if dx2 is None else torch.cat([dx1, dx2], -1)
    assert rel(got.float(), want_dx) < 5e-3
    assert rel(dg - 3.0, 2 * g_.grad) < 2e-3 and rel(db + 2.0, 2 * b_.grad) < 2e-3
    # frozen layer: no affine gradients, statistics handed in from the forward pass
    ws = torch.zeros(ops.gn_ws_floats(B, groups), device="cuda")
    ops.groupnorm_stats(x1, x2, ws, B=B, HW=HW, groups=groups)
    dx1b = torch.full_like(x1, float("nan"))
    dx2b = None if x2 is None else torch.full_like(x2, float("nan"))
    ops.groupnorm_bwd(x1, x2, dy, gamma, beta, dx1b, dx2b, None, B=B, HW=HW, groups=groups, eps=eps, silu=silu, stats=ws)
    gotb = dx1b if dx2b is None else torch.cat([dx1b, dx2b], -1)
    assert rel(gotb.float(), xcat.grad) < 5e-3


@pytest.mark.parametrize("rows,C", [(8192, 320), (2048, 640), (513, 1280), (100, 128), (64, 64)])
def test_layernorm_backward_vs_autograd(ops, rows, C):
    x, dy, dres = randn(rows, C, seed=1, scale=2.0) + 0.3, randn(rows, C, seed=2), randn(rows, C, seed=3)
    gamma = 1 + 0.2 * randn(C, seed=4, dtype=torch.float32)
    xf = x.float().requires_grad_(True)
    F.layer_norm(xf, (C,), gamma, torch.zeros_like(gamma), 1e-5).backward(dy.float())
    dx = torch.full_like(x, float("nan"))
    ops.layernorm_bwd(x, dy, gamma, dx, 1e-5, dres=dres)
    assert rel(dx.float(), xf.grad + dres.float()) < 4e-3
    ops.layernorm_bwd(x, dy, gamma, dx, 1e-5)
    assert rel(dx.float(), xf.grad) < 4e-3


@pytest.mark.parametrize("rows,C", [(4096, 1280), (1000, 2560), (77, 256)])
def test_geglu_forward_backward_vs_autograd(ops, rows, C):
    proj, d_out = randn(rows, 2 * C, seed=1, scale=1.5), randn(rows, C, seed=2)
    pf = proj.float().requires_grad_(True)
    h, gate = pf.chunk(2, -1)
    ref = h * F.gelu(gate)
    ref.backward(d_out.float())
    out, d_proj = torch.full((rows, C), float("nan"), device="cuda", dtype=bf16), torch.full((rows, 2 * C), float("nan"), device="cuda", dtype=bf16)
    ops.geglu(proj, out=out, d_out=d_out, d_proj=d_proj)
    assert rel(out.float(), ref) < 4e-3 and rel(d_proj.float(), pf.grad) < 4e-3
    out2 = torch.full_like(out, float("nan"))
    ops.geglu(proj, out=out2)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,H,W,Cin", [(2, 16, 16, 64), (3, 9, 7, 320), (2, 64, 64, 320)])
def test_conv_out_backward_vs_autograd(ops, B, H, W, Cin):
    Cout = 4
    x = randn(B, Cin, H, W, seed=1, dtype=torch.float32).requires_grad_(True)
    w = randn(Cout, Cin, 3, 3, seed=2, dtype=torch.float32) * 0.05
    dy = randn(B, Cout, H, W, seed=3, dtype=torch.float32)
    F.conv2d(x, w, None, padding=1).backward(dy)
    dx = torch.full((B, H * W, Cin), float("nan"), device="cuda", dtype=bf16)
    ops.conv_out_bwd(dy.contiguous(), w.permute(0, 2, 3, 1).contiguous(), dx, B=B, H=H, W=W)
    assert rel(dx.float().view(B, H, W, Cin).permute(0, 3, 1, 2), x.grad) < 4e-3


@pytest.mark.parametrize("Cout,Cin,k", [(320, 320, 3), (640, 960, 3), (320, 640, 1), (80, 96, 3), (1280, 2560, 1)])
def test_dgrad_repack_equals_the_torch_packing(ops, Cout, Cin, k):
    w = randn(Cout, Cin, k, k, seed=1, dtype=torch.float32)
    wp = ops.pack_conv_weight(w)                                   # [Cout, k*k*Cin] bf16
    want = ops.pack_conv_dgrad_weight(w)                           # [Cin, k*k*Cout] bf16 (rounding commutes with the permutation)
    wd = torch.full((Cin, k * k * Cout), float("nan"), device="cuda", dtype=bf16)
    ops.dgrad_repack(wp, wd, k)
    assert torch.equal(wd, want)


@pytest.mark.parametrize("B,H,W,C", [(2, 8, 8, 1280), (3, 5, 7, 64), (2, 32, 32, 640)])
def test_sumpool2x2_is_the_adjoint_of_nearest_upsampling(ops, B, H, W, C):
    du = randn(B, 4 * H * W, C, seed=2)
    dx = torch.full((B, H * W, C), float("nan"), device="cuda", dtype=bf16)
    ops.sumpool2x2(du, dx, B=B, H=H, W=W)
    want = du.float().view(B, H, 2, W, 2, C).sum((2, 4)).reshape(B, H * W, C)
    assert rel(dx.float(), want) < 4e-3
