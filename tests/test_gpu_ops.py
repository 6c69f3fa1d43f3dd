"""Kernel-level parity: every C-ABI kernel against a plain PyTorch fp32 evaluation of the same op on the same
(bf16-rounded) inputs.  Tolerances are rel-L2 in fp32: bf16 output rounding alone is ~1.5e-3."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from mirrorfusion_b200 import ops as o
    o.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return o


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def randn(*shape, seed=0, scale=1.0, dtype=bf16):
    return (torch.randn(*shape, generator=_g(seed), device="cuda") * scale).to(dtype)


def nhwc_to_nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------------------------------------ implicit GEMM
@pytest.mark.parametrize("M,K,N", [(512, 320, 320), (154, 768, 320), (8192, 320, 960), (100, 64, 8), (128, 1280, 1280)])
def test_linear(ops, M, K, N):
    x = randn(M, K, seed=1)
    w = randn(N, K, seed=2, scale=K ** -0.5)
    b = randn(N, seed=3, dtype=torch.float32)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=bf16)
    ops.linear_plan(x, w, out, bias=b).run()
    ref = x.float() @ w.float().t() + b
    assert rel(out.float(), ref) < 4e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 128, 320), (4, 8, 8, 64, 128), (2, 64, 64, 320, 320),
                                            (3, 8, 8, 64, 160), (1, 24, 24, 64, 160), (2, 12, 12, 128, 128)])
def test_conv3x3_full_epilogue(ops, B, H, W, Cin, Cout):
    x = randn(B, H, W, Cin, seed=1)
    w = randn(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5)
    bias = randn(Cout, seed=3, dtype=torch.float32)
    rowbias = randn(B, Cout + 64, seed=4, dtype=torch.float32)   # a slice of a wider table
    alpha = torch.tensor([0.75], device="cuda")
    r1 = randn(B, H, W, Cout, seed=5)
    r2 = randn(B, H, W, Cout, seed=6)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=bf16)
    wp = ops.pack_conv_weight(w.float())
    # rowbias is a [B, ld] table with ld > Cout: the kernel reads the first Cout columns of each row
    plan = ops.ConvPlan(x, wp, out, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3, bias=bias,
                        rowbias=rowbias, rowbias_ld=Cout + 64, alpha=alpha, res1=r1, res2=r2)
    plan.run()
    ref = F.conv2d(nhwc_to_nchw(x), w.float(), None, padding=1)
    ref = (ref + bias.view(1, -1, 1, 1) + rowbias[:, :Cout, None, None]) * 0.75 + nhwc_to_nchw(r1) + nhwc_to_nchw(r2)
    assert rel(nhwc_to_nchw(out), ref) < 4e-3


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 64), (2, 16, 16, 128), (1, 64, 64, 320), (2, 24, 24, 64)])
def test_conv3x3_stride2(ops, B, H, W, C):
    x = randn(B, H, W, C, seed=1)
    w = randn(C, C, 3, 3, seed=2, scale=(9 * C) ** -0.5)
    bias = randn(C, seed=3, dtype=torch.float32)
    tap = randn(B, H // 2, W // 2, C, seed=4)
    out = torch.full((B, H // 2, W // 2, C), float("nan"), device="cuda", dtype=bf16)
    ops.ConvPlan(x, ops.pack_conv_weight(w.float()), out, B=B, H=H, W=W, Cin=C, Cout=C, ksize=3, stride=2, bias=bias,
                 res2=tap).run()
    ref = F.conv2d(nhwc_to_nchw(x), w.float(), bias, stride=2, padding=1) + nhwc_to_nchw(tap)
    assert rel(nhwc_to_nchw(out), ref) < 4e-3


@pytest.mark.parametrize("B,H,W,C,with_extra", [(2, 8, 8, 128, False), (2, 16, 16, 64, True), (1, 32, 32, 320, True), (3, 12, 12, 64, False)])
def test_upsample_fused_conv(ops, B, H, W, C, with_extra):
    """Upsample2D = nearest x2 + conv3x3 as four sub-pixel 2x2 convs over the low-res input (no upsampled tensor),
    with the tap add and an extra 1x1 K-segment living at the OUTPUT resolution."""
    x = randn(B, H, W, C, seed=1)
    w = randn(C, C, 3, 3, seed=2, scale=(9 * C) ** -0.5)
    bias = randn(C, seed=3, dtype=torch.float32)
    tap = randn(B, 2 * H, 2 * W, C, seed=4)
    ex = randn(B, 2 * H, 2 * W, 64, seed=5) if with_extra else None
    wz = randn(C, 64, seed=6, scale=0.1) if with_extra else None
    out = torch.full((B, 2 * H, 2 * W, C), float("nan"), device="cuda", dtype=bf16)
    wp = ops.pack_upconv_weight(w.float(), extras=[wz.float()] if with_extra else [])
    plan = ops.ConvPlan(x, wp, out, B=B, H=H, W=W, Cin=C, Cout=C, ksize=3, up2x=True, bias=bias, res2=tap,
                        extras=[ex] if with_extra else [])
    assert plan.launches == 4
    plan.run()
    up = F.interpolate(nhwc_to_nchw(x), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), bias, padding=1) + nhwc_to_nchw(tap)
    if with_extra:
        ref = ref + F.conv2d(nhwc_to_nchw(ex), wz.float()[:, :, None, None])
    assert rel(nhwc_to_nchw(out), ref) < 5e-3
    assert abs(plan.flops - 2.0 * B * 4 * H * W * C * (9 * C + (64 if with_extra else 0))) < 1.0   # algorithmic FLOPs


def test_conv3x3_with_shortcut_segments(ops):
    # ResnetBlock2D tail on an up block: conv2(h) + conv_shortcut(cat([x, skip])) fused in one accumulator
    B, H, W, Cmid, Ca, Cb = 2, 16, 16, 128, 128, 64
    h = randn(B, H, W, Cmid, seed=1)
    xa = randn(B, H, W, Ca, seed=2)
    xb = randn(B, H, W, Cb, seed=3)
    w2 = randn(Cmid, Cmid, 3, 3, seed=4, scale=(9 * Cmid) ** -0.5)
    ws = randn(Cmid, Ca + Cb, 1, 1, seed=5, scale=(Ca + Cb) ** -0.5)
    bias = randn(Cmid, seed=6, dtype=torch.float32)
    out = torch.full((B, H, W, Cmid), float("nan"), device="cuda", dtype=bf16)
    wp = ops.pack_conv_weight(w2.float(), extras=[ws.float()[:, :Ca, 0, 0], ws.float()[:, Ca:, 0, 0]])
    ops.ConvPlan(h, wp, out, B=B, H=H, W=W, Cin=Cmid, Cout=Cmid, ksize=3, extras=[xa, xb], bias=bias).run()
    ref = F.conv2d(nhwc_to_nchw(h), w2.float(), bias, padding=1) + F.conv2d(
        torch.cat([nhwc_to_nchw(xa), nhwc_to_nchw(xb)], 1), ws.float())
    assert rel(nhwc_to_nchw(out), ref) < 4e-3


def test_geglu_linear(ops):
    M, Cc = 384, 320
    x = randn(M, Cc, seed=1)
    w = randn(8 * Cc, Cc, seed=2, scale=Cc ** -0.5)
    b = randn(8 * Cc, seed=3, dtype=torch.float32)
    wp, bp = ops.pack_geglu(w.float(), b)
    out = torch.full((M, 4 * Cc), float("nan"), device="cuda", dtype=bf16)
    ops.linear_plan(x, wp, out, bias=bp, geglu=True).run()
    y = x.float() @ w.float().t() + b
    val, gate = y.chunk(2, -1)
    assert rel(out.float(), val * F.gelu(gate)) < 4e-3


def test_conv_block_n_variants_agree(ops):
    B, H, W, Cin, Cout = 2, 16, 16, 64, 640
    x = randn(B, H, W, Cin, seed=1)
    wp = ops.pack_conv_weight(randn(Cout, Cin, 3, 3, seed=2, scale=0.05).float())
    outs = []
    for bn in (64, 80, 128, 160):
        o = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=bf16)
        ops.ConvPlan(x, wp, o, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3, block_n=bn).run()
        outs.append(o)
    for o in outs[1:]:
        assert torch.equal(outs[0], o)    # same K order, same fp32 accumulation -> bit identical


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 128, 320), (3, 16, 16, 64, 128), (2, 64, 64, 320, 640)])
def test_igemm_cluster_modes_agree(ops, B, H, W, Cin, Cout):
    """igemm_mode 1 = independent CTAs, 2 = CTA pair + TMA-multicast weight tile, 3 = CTA pair + tcgen05.mma.cta_group::2
    (256-row UMMA tile, accumulators in both SMs' TMEM).  Same K order and fp32 accumulation -> bit-identical outputs
    (odd M-tile counts exercise the all-out-of-bounds partner tile)."""
    x = randn(B, H, W, Cin, seed=1)
    wp = ops.pack_conv_weight(randn(Cout, Cin, 3, 3, seed=2, scale=0.05).float())
    bias = randn(Cout, seed=3, dtype=torch.float32)
    r1 = randn(B, H, W, Cout, seed=4)
    outs = []
    for mode in (1, 2, 3):
        o = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=bf16)
        ops.ConvPlan(x, wp, o, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3, bias=bias, res1=r1, igemm_mode=mode).run()
        outs.append(o)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,variant", [
    (16, 8, 8, 1280, 1280, 1, "plain"), (16, 8, 8, 1280, 1280, 1, "res1+rowbias"), (16, 8, 8, 320, 1280, 1, "odd K blocks"),
    (16, 16, 16, 1280, 1280, 2, "stride 2"), (16, 8, 8, 1280, 1280, 1, "shortcut segments"), (4, 8, 8, 640, 640, 1, "plain"),
    (16, 8, 8, 1280, 1280, 1, "up2x"), (3, 8, 8, 128, 256, 1, "ragged M")])
def test_igemm_split_k_pair(ops, B, H, W, Cin, Cout, stride, variant):
    """igemm_mode 4: few-tile launches (the 8x8 level: M = 64 * B) on the wide tile with TWO CTAs per tile, each half of the K
    blocks, rank 1's fp32 partial accumulator handed to rank 0 through distributed shared memory.  Against torch AND against the
    independent-CTA plan (fp32 sum split in two halves: equal to rounding of the bf16 output)."""
    x = randn(B, H, W, Cin, seed=1)
    w = randn(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5)
    bias = randn(Cout, seed=3, dtype=torch.float32)
    Ho, Wo = (2 * H, 2 * W) if variant == "up2x" else (H // stride, W // stride)
    kw = dict(B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3, bias=bias)
    ref = None
    if variant == "up2x":
        wp = ops.pack_upconv_weight(w.float())
        kw["up2x"] = True
        ref = F.conv2d(F.interpolate(nhwc_to_nchw(x), scale_factor=2.0, mode="nearest"), w.float(), bias, padding=1)
    elif variant == "shortcut segments":
        xa, xb = randn(B, H, W, 1280, seed=7), randn(B, H, W, 640, seed=8)
        ws = randn(Cout, 1920, 1, 1, seed=9, scale=1920 ** -0.5)
        wp = ops.pack_conv_weight(w.float(), extras=[ws.float()[:, :1280, 0, 0], ws.float()[:, 1280:, 0, 0]])
        kw["extras"] = [xa, xb]
        ref = F.conv2d(nhwc_to_nchw(x), w.float(), bias, padding=1) + F.conv2d(torch.cat([nhwc_to_nchw(xa), nhwc_to_nchw(xb)], 1), ws.float())
    else:
        wp = ops.pack_conv_weight(w.float())
        kw["stride"] = stride
        ref = F.conv2d(nhwc_to_nchw(x), w.float(), bias, stride=stride, padding=1)
    if variant == "res1+rowbias":
        r1 = randn(B, Ho, Wo, Cout, seed=5)
        rowbias = randn(B, Cout, seed=4, dtype=torch.float32)
        kw.update(res1=r1, rowbias=rowbias, rowbias_ld=Cout)
        ref = ref + rowbias[:, :, None, None] + nhwc_to_nchw(r1)
    outs = []
    for mode in (4, 1):
        o = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda", dtype=bf16)
        plan = ops.ConvPlan(x, wp, o, igemm_mode=mode, **kw)
        assert plan.mode == (3 if mode == 4 else 0), (mode, plan.mode)
        plan.run()
        torch.cuda.synchronize()
        outs.append(o)
    assert rel(nhwc_to_nchw(outs[0]), ref) < 4e-3
    assert rel(outs[0].float(), outs[1].float()) < 2e-3          # differ only where the two-half sum rounds to another bf16 value
    # deterministic: the hand-over adds the halves in a fixed order
    o2 = torch.full_like(outs[0], float("nan"))
    ops.ConvPlan(x, wp, o2, igemm_mode=4, **kw).run()
    assert torch.equal(o2, outs[0])


def test_igemm_split_k_falls_back_when_the_tile_count_is_large(ops):
    x = randn(8, 32, 32, 128, seed=1)                      # 64 M tiles x 2 N tiles: more tiles than CTA pairs
    wp = ops.pack_conv_weight(randn(320, 128, 3, 3, seed=2, scale=0.03).float())
    o = torch.empty(8, 32, 32, 320, device="cuda", dtype=bf16)
    assert ops.ConvPlan(x, wp, o, B=8, H=32, W=32, Cin=128, Cout=320, ksize=3, igemm_mode=4).mode == 0


def test_persistent_tile_loop_many_tiles(ops):
    # far more tiles than SMs: every CTA walks several tiles through both TMEM accumulator slots
    M, K, N = 40000, 128, 640
    x = randn(M, K, seed=1)
    w = randn(N, K, seed=2, scale=K ** -0.5)
    r1 = randn(M, N, seed=3)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=bf16)
    ops.linear_plan(x, w, out, res1=r1).run()
    assert rel(out.float(), x.float() @ w.float().t() + r1.float()) < 4e-3


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("B,HW,C1,C2,groups,eps,silu", [
    (2, 4096, 320, 0, 32, 1e-5, True), (2, 256, 1280, 640, 32, 1e-5, True), (2, 1024, 640, 320, 32, 1e-5, True),
    (3, 64, 1280, 1280, 32, 1e-5, True), (2, 1024, 640, 0, 32, 1e-6, False), (2, 256, 64, 0, 8, 1e-5, True),
    # single-launch path of the small feature maps: group pairs (cpg 20 / 60), one group per CTA (cpg 40 / 80)
    (2, 256, 640, 0, 32, 1e-5, True), (3, 64, 1280, 0, 32, 1e-5, True), (2, 256, 1280, 1280, 32, 1e-5, False),
    (16, 4096, 320, 0, 32, 1e-5, True)])
def test_groupnorm(ops, B, HW, C1, C2, groups, eps, silu):
    x1 = randn(B, HW, C1, seed=1) + 0.5
    x2 = randn(B, HW, C2, seed=2, scale=2.0) if C2 else None
    Cc = C1 + C2
    gamma = 1 + 0.1 * randn(Cc, seed=3, dtype=torch.float32)
    beta = 0.1 * randn(Cc, seed=4, dtype=torch.float32)
    out = torch.full((B, HW, Cc), float("nan"), device="cuda", dtype=bf16)
    ws = torch.zeros(ops.gn_ws_floats(B, groups), device="cuda")
    ops.groupnorm(x1, x2, gamma, beta, out, ws, B=B, HW=HW, groups=groups, eps=eps, silu=silu)
    first = out.clone()
    ops.groupnorm(x1, x2, gamma, beta, out, ws, B=B, HW=HW, groups=groups, eps=eps, silu=silu)   # counters re-armed, bitwise repeatable
    assert torch.equal(first, out)
    xin = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)
    ref = F.group_norm(xin.transpose(1, 2), groups, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    assert rel(out.float(), ref.transpose(1, 2)) < 4e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,up", [(2, 32, 32, 64, 320, False), (4, 8, 8, 64, 640, False), (2, 12, 12, 64, 160, False),
                                               (2, 8, 8, 64, 320, True), (3, 16, 16, 128, 1280, False)])
def test_groupnorm_with_statistics_from_the_conv_epilogue(ops, B, H, W, Cin, Cout, up):
    """The igemm epilogue emits per-(image, tile, channel) sum / sum-of-squares of the values it stores; GroupNorm then
    only runs a tiny fixed-order reduction instead of re-reading the tensor.  Must equal the two-pass GroupNorm bit for
    bit up to fp32 summation order, and the second source of a concat may come from another conv."""
    groups = 32
    Ho, Wo = (2 * H, 2 * W) if up else (H, W)
    x = randn(B, H, W, Cin, seed=1)
    w = randn(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5)
    bias = randn(Cout, seed=3, dtype=torch.float32)
    res = randn(B, Ho, Wo, Cout, seed=4)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda", dtype=bf16)
    wp = ops.pack_upconv_weight(w.float()) if up else ops.pack_conv_weight(w.float())
    plan = ops.ConvPlan(x, wp, out, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3, bias=bias, res1=None if up else res,
                        res2=res if up else None, up2x=up)
    st = plan.enable_output_stats()
    if st is None:       # tiles of this geometry straddle images: the engine then falls back to the statistics kernel
        assert (B, H) == (2, 12)
        return
    plan.run()
    part, tiles = st
    got = part.view(B, tiles, Cout, 2).sum(1)                       # [B, C, 2]
    o = out.float().view(B, Ho * Wo, Cout)
    assert rel(got[..., 0], o.sum(1)) < 1e-4 and rel(got[..., 1], (o * o).sum(1)) < 1e-5
    # GroupNorm + SiLU over cat([out, y2]) with y2's statistics computed the classic way for the first call ...
    gamma = 1 + 0.1 * randn(Cout, seed=5, dtype=torch.float32)
    beta = 0.1 * randn(Cout, seed=6, dtype=torch.float32)
    ws = torch.zeros(ops.gn_ws_floats(B, groups), device="cuda")
    y_fused = torch.full((B, Ho * Wo, Cout), float("nan"), device="cuda", dtype=bf16)
    y_plain = torch.full_like(y_fused, float("nan"))
    ops.groupnorm(out.view(B, Ho * Wo, Cout), None, gamma, beta, y_fused, ws, B=B, HW=Ho * Wo, groups=groups, eps=1e-5, silu=True,
                  part1=st)
    ops.groupnorm(out.view(B, Ho * Wo, Cout), None, gamma, beta, y_plain, ws, B=B, HW=Ho * Wo, groups=groups, eps=1e-5, silu=True)
    ref = F.silu(F.group_norm(o.transpose(1, 2), groups, gamma, beta, 1e-5)).transpose(1, 2)
    assert rel(y_fused.float(), ref) < 4e-3 and rel(y_fused.float(), y_plain.float()) < 2e-3


@pytest.mark.parametrize("rows,C", [(8192, 320), (2048, 640), (513, 1280), (1001, 320), (515, 640), (64, 64), (100, 128)])
def test_layernorm(ops, rows, C):
    x = randn(rows, C, seed=1) * 2 + 0.3
    gamma = 1 + 0.1 * randn(C, seed=2, dtype=torch.float32)
    beta = 0.1 * randn(C, seed=3, dtype=torch.float32)
    out = torch.full((rows, C), float("nan"), device="cuda", dtype=bf16)
    ops.layernorm(x, gamma, beta, out)
    assert rel(out.float(), F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)) < 4e-3


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,heads,d,Tq,Tk", [
    (2, 8, 40, 4096, 4096), (2, 8, 80, 1024, 1024), (2, 8, 160, 256, 256), (2, 8, 160, 64, 64),
    (2, 8, 40, 4096, 77), (2, 8, 80, 1024, 77), (2, 8, 160, 256, 77), (2, 8, 160, 64, 77),
    (2, 2, 32, 256, 256), (2, 2, 64, 64, 64), (2, 2, 64, 16, 77), (1, 8, 40, 200, 333),
    (1, 8, 40, 9216, 9216), (1, 8, 40, 9216, 77), (1, 8, 80, 2304, 2304),       # config 5: 96x96 latents (768x768 images)
    (1, 8, 40, 16384, 16384)])                                                    # 128x128 latents (1024x1024 images)
def test_attention(ops, B, heads, d, Tq, Tk):
    Cc = heads * d
    q = randn(B, Tq, Cc, seed=1)
    k = randn(B, Tk, Cc, seed=2)
    v = randn(B, Tk, Cc, seed=3)
    out = torch.full((B, Tq, Cc), float("nan"), device="cuda", dtype=bf16)
    ops.attention(q, k, v, out, B=B, heads=heads, head_dim=d, Tq=Tq, Tk=Tk)
    qh = q.float().view(B, Tq, heads, d).transpose(1, 2)
    kh = k.float().view(B, Tk, heads, d).transpose(1, 2)
    vh = v.float().view(B, Tk, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * d ** -0.5
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Tq, Cc)
    assert rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("d,T", [(40, 1024), (80, 512), (160, 256)])
def test_attention_rescale_path(ops, d, T):
    # keys grow along the sequence, so the running row maximum jumps by more than the lazy-rescale threshold (2^8)
    # in later key tiles and the O accumulator in TMEM is rescaled many times (randn inputs almost never do that)
    B, heads = 2, 4
    Cc = heads * d
    q = randn(B, T, Cc, seed=11) * 3.0
    ramp = torch.linspace(0.2, 6.0, T, device="cuda").view(1, T, 1)
    k = (randn(B, T, Cc, seed=12).float() * ramp).to(bf16)
    v = randn(B, T, Cc, seed=13)
    out = torch.full((B, T, Cc), float("nan"), device="cuda", dtype=bf16)
    ops.attention(q, k, v, out, B=B, heads=heads, head_dim=d, Tq=T, Tk=T)
    qh, kh, vh = [t.float().view(B, T, heads, d).transpose(1, 2) for t in (q, k, v)]
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1) @ vh).transpose(1, 2).reshape(B, T, Cc)
    assert rel(out.float(), ref) < 8e-3


def test_attention_fused_qkv_layout(ops):
    # q/k/v read straight out of a fused [B, T, 3C] projection buffer via leading dimensions
    B, heads, d, T = 2, 8, 40, 512
    Cc = heads * d
    qkv = randn(B, T, 3 * Cc, seed=7)
    out = torch.empty(B, T, Cc, device="cuda", dtype=bf16)
    ops.attention(qkv, qkv[:, :, Cc:], qkv[:, :, 2 * Cc:], out, B=B, heads=heads, head_dim=d, Tq=T, Tk=T, ldq=3 * Cc, ldk=3 * Cc,
                  ldv=3 * Cc)
    q, k, v = [t.float().view(B, T, heads, d).transpose(1, 2) for t in qkv.split(Cc, -1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v).transpose(1, 2).reshape(B, T, Cc)
    assert rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("d", [40, 80])
def test_attention_v_ride_along_columns_never_stored(ops, d):
    # the V tile is fetched in 64-column boxes, so for d = 40 / 80 columns past a head's own ride along into accumulator
    # columns [d, dpad); NaNs planted in the row padding after the last head must not reach any stored output
    B, heads, T = 1, 8, 300
    Cc, ldv = heads * d, heads * d + 24
    q, k = randn(B, T, Cc, seed=21), randn(B, T, Cc, seed=22)
    vbuf = torch.full((B, T, ldv), float("nan"), device="cuda", dtype=bf16)
    vbuf[:, :, :Cc] = randn(B, T, Cc, seed=23)
    out = torch.full((B, T, Cc), float("nan"), device="cuda", dtype=bf16)
    ops.attention(q, k, vbuf, out, B=B, heads=heads, head_dim=d, Tq=T, Tk=T, ldv=ldv)
    qh, kh, vh = [t.float().view(B, T, heads, d).transpose(1, 2) for t in (q, k, vbuf[:, :, :Cc])]
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1) @ vh).transpose(1, 2).reshape(B, T, Cc)
    assert bool(torch.isfinite(out).all())
    assert rel(out.float(), ref) < 6e-3


# ------------------------------------------------------------------------------------------------ boundary / misc
@pytest.mark.parametrize("Ca,Cb,H,W", [(4, 0, 64, 64), (4, 6, 64, 64), (4, 6, 20, 12)])
def test_conv_in(ops, Ca, Cb, H, W):
    B, Cout = 2, 320
    xa = randn(B, Ca, H, W, seed=1, dtype=torch.float32)
    xb = randn(B, Cb, H, W, seed=2, dtype=torch.float32) if Cb else None
    w = randn(Cout, Ca + Cb, 3, 3, seed=3, scale=0.1, dtype=torch.float32)
    b = randn(Cout, seed=4, dtype=torch.float32)
    tap = randn(B, H, W, Cout, seed=5)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=bf16)
    post = torch.empty_like(out)
    ops.conv_in(xa, xb, w.permute(2, 3, 1, 0).contiguous(), b, out, tap, post)
    xin = xa if xb is None else torch.cat([xa, xb], 1)
    ref = F.conv2d(xin, w, b, padding=1)
    assert rel(nhwc_to_nchw(out), ref) < 4e-3
    assert rel(nhwc_to_nchw(post), ref + nhwc_to_nchw(tap)) < 4e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 4), (16, 64, 64, 320, 4), (3, 5, 7, 64, 4), (1, 16, 16, 72, 3),
                                            (2, 24, 40, 64, 4)])
def test_conv_out(ops, B, H, W, Cin, Cout):
    # sizes whose pixel count is / is not a multiple of the 32-pixel CTA step, channel-vector counts that do / do not
    # fill the 8 lanes of a pixel group (Cin 72 -> 9 vectors), fewer than 4 outputs, more pixel groups than resident CTAs
    x = randn(B, H, W, Cin, seed=1)
    w = randn(Cout, Cin, 3, 3, seed=2, scale=0.02, dtype=torch.float32)
    b = randn(Cout, seed=3, dtype=torch.float32)
    out = torch.full((B, Cout, H, W), float("nan"), device="cuda")
    ops.conv_out(x, w.permute(0, 2, 3, 1).contiguous(), b, out, B=B, H=H, W=W)
    assert rel(out, F.conv2d(nhwc_to_nchw(x), w, b, padding=1)) < 1e-4


def test_upsample_and_layout(ops):
    B, H, W, Cc = 2, 8, 8, 128
    x = randn(B, H, W, Cc, seed=1)
    up = torch.empty(B, 2 * H, 2 * W, Cc, device="cuda", dtype=bf16)
    ops.upsample2x(x, up, B=B, H=H, W=W)
    ref = F.interpolate(nhwc_to_nchw(x), scale_factor=2.0, mode="nearest")
    assert torch.equal(nhwc_to_nchw(up), ref)
    f = randn(B, 5, 7, 9, seed=2, dtype=torch.float32)
    nh = torch.empty(B, 7, 9, 5, device="cuda", dtype=bf16)
    ops.nchw_to_nhwc(f, nh)
    assert torch.equal(nh, f.permute(0, 2, 3, 1).to(bf16))
    back = torch.empty(B, 5, 7, 9, device="cuda")
    ops.nhwc_to_nchw(nh, back)
    assert torch.equal(back, nh.float().permute(0, 3, 1, 2))


def test_timestep_path(ops):
    t = torch.tensor([999.0, 500.0, 1.0], device="cuda")
    emb = torch.empty(3, 320, device="cuda")
    ops.timestep_sinusoid(t, emb)
    half = 160
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    a = t[:, None] * freq[None]
    assert torch.allclose(emb, torch.cat([a.cos(), a.sin()], -1), atol=2e-4)
    w = randn(1280, 320, seed=1, scale=0.05)
    b = randn(1280, seed=2, dtype=torch.float32)
    y = torch.empty(3, 1280, device="cuda")
    ops.linear_small(emb, w, b, y, act_in=False, act_out=True)
    assert rel(y, F.silu(emb @ w.float().t() + b)) < 1e-5
    y2 = torch.empty(3, 1280, device="cuda")
    ops.linear_small(emb, w, b, y2, act_in=True, act_out=False)
    assert rel(y2, F.silu(emb) @ w.float().t() + b) < 1e-5


def test_cfg_sched_kernel(ops):
    Bi, n = 3, 4 * 64 * 64
    eps = randn(2 * Bi, n, seed=1, dtype=torch.float32)
    x = randn(Bi, n, seed=2, dtype=torch.float32)
    last = randn(Bi, n, seed=3, dtype=torch.float32)
    m0 = randn(Bi, n, seed=4, dtype=torch.float32)
    m1 = randn(Bi, n, seed=5, dtype=torch.float32)
    coef = torch.tensor([7.5, 1.1, -0.4, 0.9, 0.2, -0.05, 0.3, 1.0, 0.8, 0.15, -0.07, 0.0], device="cuda")
    e = eps[:Bi] + 7.5 * (eps[Bi:] - eps[:Bi])
    mt = 1.1 * x - 0.4 * e
    xc = 0.9 * last + 0.2 * m0 - 0.05 * m1 + 0.3 * mt
    xn = 0.8 * xc + 0.15 * mt - 0.07 * m0
    m0_old = m0.clone()
    ops.cfg_sched_step(eps[:Bi], eps[Bi:], x, last, m0, m1, coef)
    assert rel(x, xn) < 1e-6 and rel(last, xc) < 1e-6 and rel(m0, mt) < 1e-6 and torch.equal(m1, m0_old)


@pytest.mark.parametrize("B,Cin,with_res,with_extra", [(16, 1280, True, False), (16, 2560, False, True), (2, 1280, False, False)])
def test_conv3x3_8x8_level_long_k(ops, B, Cin, with_res, with_extra):
    # the 8x8 latent level: 8 M tiles (or 1), K = 11520 .. 24320 with a residual or an extra 1x1 segment; run twice (plans are
    # replayed from a CUDA graph: nothing may depend on state left by the previous launch)
    H = W = 8
    Cout = 1280
    x = randn(B, H, W, Cin, seed=31)
    w = randn(Cout, Cin, 3, 3, seed=32, dtype=torch.float32) * (Cin * 9) ** -0.5
    bias = 0.1 * randn(Cout, seed=33, dtype=torch.float32)
    res = randn(B, H, W, Cout, seed=34) if with_res else None
    ex = randn(B, H, W, 1280, seed=35) if with_extra else None
    wex = (randn(Cout, 1280, seed=36, dtype=torch.float32) * 1280 ** -0.5) if with_extra else None
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=bf16)
    plan = ops.ConvPlan(x, ops.pack_conv_weight(w, extras=[wex] if with_extra else []), out, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3,
                        extras=[ex] if with_extra else [], bias=bias, res1=res)
    ref = F.conv2d(nhwc_to_nchw(x), w.to(bf16).float(), bias, padding=1)
    if with_extra:
        ref = ref + F.conv2d(nhwc_to_nchw(ex), wex.to(bf16).float()[:, :, None, None])
    if with_res:
        ref = ref + nhwc_to_nchw(res)
    for _ in range(2):
        out.fill_(float("nan"))
        plan.run()
        assert rel(nhwc_to_nchw(out), ref) < 5e-3
