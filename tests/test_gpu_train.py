"""Fine-tune step glue on the GPU (csrc/train.cu through the C ABI): against the reference's own vectors
(tests/golden/train_glue.npz), the oracle (oracle/train_oracle.py) and torch's AdamW / clip_grad_norm_ / conv autograd —
the functions E/train_brushnet_mirror.py:1404-1466 calls.  Tolerances are written at each assert."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import train_oracle as T

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


def _randn(*shape, seed=0, dtype=torch.float32, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_add_noise_and_velocity_vs_reference_golden(ops, golden_dir):
    from mirrorfusion_b200.train import NoiseSchedule
    g = np.load(os.path.join(golden_dir, "train_glue.npz"))
    ns = NoiseSchedule("cuda")
    x0, noise, t = cu(g["x0"]), cu(g["noise"]), cu(g["t"])
    noisy = ns.add_noise(x0, noise, t)
    vel = ns.get_velocity(x0, noise, t)
    # fp32 elementwise: sqrt / fma contraction differences only
    assert torch.allclose(noisy.cpu(), torch.from_numpy(g["noisy"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(vel.cpu(), torch.from_numpy(g["velocity"]), rtol=1e-6, atol=1e-6)
    both_n, both_v = torch.empty_like(x0), torch.empty_like(x0)
    ops.add_noise(x0, noise, t, ns.alphas_cumprod, noisy=both_n, velocity=both_v)
    assert torch.equal(both_n, noisy) and torch.equal(both_v, vel)


def test_add_noise_large_ragged(ops):
    from mirrorfusion_b200.train import NoiseSchedule
    ns = NoiseSchedule("cuda")
    B, n = 33, 4 * 64 * 64 + 3
    x0, noise = _randn(B, n, seed=1), _randn(B, n, seed=2)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(3))
    got = ns.add_noise(x0, noise, t.cuda())
    ref = T.add_noise(x0.cpu().numpy(), noise.cpu().numpy(), t.numpy(), ns.acp_host)
    assert torch.allclose(got.cpu(), torch.from_numpy(ref), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["plain", "snr5"])
def test_mse_loss_and_gradient_vs_reference_golden(ops, golden_dir, name):
    from mirrorfusion_b200.train import TrainLoss
    g = np.load(os.path.join(golden_dir, "train_glue.npz"))
    pred, target = cu(g["pred"]), cu(g["noise"])
    w = None if name == "plain" else cu(g["w_snr5"])
    L = TrainLoss(pred.shape[0], "cuda")
    grad = torch.full_like(pred, float("nan"))
    loss = L(pred, target, weights=w, grad=grad)
    assert abs(loss.item() - float(g[f"loss_{name}"])) < 2e-6 * abs(float(g[f"loss_{name}"]))
    assert torch.allclose(grad.cpu(), torch.from_numpy(g[f"grad_{name}"]), rtol=2e-5, atol=1e-9)
    _, per, _ = T.mse_loss(g["pred"], g["noise"], None)
    assert np.allclose(L.per_sample.cpu().numpy(), per, rtol=2e-6)


def test_mse_loss_config4_shape_is_deterministic(ops):
    # batch 32 per GPU, latents 4x64x64; two runs bit-identical (fixed-order reduction), value vs float64 oracle
    from mirrorfusion_b200.train import TrainLoss
    B, n = 32, 4 * 64 * 64
    pred, target = _randn(B, n, seed=4), _randn(B, n, seed=5)
    L = TrainLoss(B, "cuda")
    grad = torch.empty_like(pred)
    a = L(pred, target, grad=grad).clone()
    b = L(pred, target, grad=grad).clone()
    assert torch.equal(a, b)
    ref, _, gref = T.mse_loss(pred.cpu().numpy(), target.cpu().numpy())
    assert abs(a.item() - ref) < 1e-6 * ref
    assert rel(grad.cpu(), torch.from_numpy(gref)) < 1e-6
    assert torch.allclose(a, F.mse_loss(pred, target), rtol=1e-5)


@pytest.mark.parametrize("n", [1, 5, 4096, 1_000_003])
def test_grad_sqnorm(ops, n):
    buf = _randn(n + 4, seed=6)
    g = buf[:n]     # the buffer start is 16-byte aligned; n itself is arbitrary
    ws = torch.zeros(ops.SQNORM_WS_FLOATS, device="cuda")
    out = torch.zeros(1, device="cuda")
    ops.grad_sqnorm(g, ws, out)
    ref = float((g.double() ** 2).sum())
    assert abs(out.item() - ref) <= 2e-6 * ref
    ops.grad_sqnorm(g, ws, out, accumulate=True)
    assert abs(out.item() - 2 * ref) <= 2e-6 * 2 * ref


@pytest.mark.parametrize("max_norm", [None, 1.0])
def test_flat_adamw_matches_torch_adamw(ops, max_norm):
    """Six optimizer steps over a multi-tensor parameter set: FlatParams + B200AdamW (one launch) against
    clip_grad_norm_ + torch.optim.AdamW on the same device (what the reference's loop runs, lines 1460-1464)."""
    from mirrorfusion_b200.train import B200AdamW, FlatParams
    shapes = {"conv.weight": (64, 32, 3, 3), "conv.bias": (64,), "norm.weight": (37,), "lin.weight": (129, 65), "scalar": ()}
    gen = torch.Generator().manual_seed(11)
    sd = {k: torch.randn(s, generator=gen) for k, s in shapes.items()}
    flat = FlatParams.from_state_dict(sd, "cuda")
    opt = B200AdamW(flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    tparams = {k: torch.nn.Parameter(v.clone().cuda()) for k, v in sd.items()}
    topt = torch.optim.AdamW(list(tparams.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    assert torch.equal(flat.w("lin.weight"), sd["lin.weight"].cuda().to(bf16))
    for step in range(6):
        scale = 5.0 if step % 2 == 0 else 1e-3          # alternate between clipped and unclipped steps
        for k, s in shapes.items():
            gk = torch.randn(s, generator=gen) * scale
            flat.g(k).copy_(gk)
            tparams[k].grad = gk.clone().cuda()
        if step == 3:
            opt.param_groups[0]["lr"] = topt.param_groups[0]["lr"] = 5e-4      # an lr-scheduler change
        if max_norm is not None:
            total = torch.nn.utils.clip_grad_norm_(list(tparams.values()), max_norm)
        topt.step()
        opt.step(max_grad_norm=max_norm)
        if max_norm is not None:
            assert abs(opt.grad_norm() - total.item()) < 1e-5 * total.item()
        for k in shapes:
            # fp32 updates, different (but fixed) operation order in the two implementations
            assert torch.allclose(flat.p(k), tparams[k].detach(), rtol=2e-6, atol=2e-7), (step, k)
    assert torch.equal(flat.work, flat.param.to(bf16))          # the working copy is the rounding of the masters
    pads = torch.ones(flat.numel, dtype=torch.bool)
    for off, n in flat.table.values():
        pads[off:off + n] = False
    assert flat.param.cpu()[pads].eq(0).all()                    # alignment padding never moves


def test_optimizer_and_lr_schedule_resume_is_bit_identical(ops):
    """save -> load -> step equals an uninterrupted run bit for bit (moments, step count -> bias correction, lr
    schedule position and base lr), like accelerator.save_state / --resume_from_checkpoint (E/train_brushnet_mirror.py:1271-1300,1488-1509)."""
    from mirrorfusion_b200.train import B200AdamW, FlatParams, LRSchedule
    shapes = {"a.weight": (64, 27), "a.bias": (64,), "b.weight": (33, 65)}
    gen = torch.Generator().manual_seed(5)
    sd = {k: torch.randn(s, generator=gen) for k, s in shapes.items()}
    grads = [{k: torch.randn(s, generator=gen) for k, s in shapes.items()} for _ in range(7)]

    def run(flat, opt, sched, lo, hi):
        for i in range(lo, hi):
            for k in shapes:
                flat.g(k).copy_(grads[i][k])
            opt.step(max_grad_norm=1.0)
            sched.step()

    mk = lambda: FlatParams.from_state_dict(sd, "cuda")
    f1 = mk()
    o1 = B200AdamW(f1, lr=1e-2)
    s1 = LRSchedule(o1, "cosine", num_warmup_steps=2, num_training_steps=10)
    run(f1, o1, s1, 0, 7)
    f2 = mk()
    o2 = B200AdamW(f2, lr=1e-2)
    s2 = LRSchedule(o2, "cosine", num_warmup_steps=2, num_training_steps=10)
    run(f2, o2, s2, 0, 4)
    ck = {"params": f2.state_dict(), "opt": o2.state_dict(), "lr": s2.state_dict()}
    assert ck["opt"]["step_count"] == 4 and ck["lr"]["last_epoch"] == 4 and ck["opt"]["exp_avg"].device.type == "cpu"
    f3 = FlatParams(shapes, "cuda")
    f3.load_state_dict(ck["params"])
    o3 = B200AdamW(f3, lr=123.0)                       # wrong on purpose: the saved groups / base lr must win
    s3 = LRSchedule(o3, "cosine")
    o3.load_state_dict(ck["opt"])
    s3.load_state_dict(ck["lr"])
    assert s3.get_last_lr() == s2.get_last_lr() and s3.base_lr == 1e-2
    run(f3, o3, s3, 4, 7)
    for a, b in ((f1.param, f3.param), (f1.exp_avg, f3.exp_avg), (f1.exp_avg_sq, f3.exp_avg_sq), (f1.work, f3.work)):
        assert torch.equal(a, b)
    assert s1.get_last_lr() == s3.get_last_lr() and o1.step_count == o3.step_count == 7
    with pytest.raises(ValueError):
        B200AdamW(FlatParams({"x": (3,)}, "cuda")).load_state_dict(ck["opt"])


def test_adamw_grad_scale_is_the_allreduce_mean(ops):
    # SUM-all-reduced gradients of W ranks with grad_scale 1/W == the mean gradient, clipping on the mean's norm
    from mirrorfusion_b200.train import B200AdamW, FlatParams
    sd = {"w": _randn(1000, seed=1).cpu()}
    g = _randn(1000, seed=2) * 3.0
    a, b = FlatParams.from_state_dict(sd, "cuda"), FlatParams.from_state_dict(sd, "cuda")
    a.grad[:1000].copy_(g * 8)
    b.grad[:1000].copy_(g)
    oa, ob = B200AdamW(a, lr=1e-2), B200AdamW(b, lr=1e-2)
    oa.step(max_grad_norm=1.0, grad_scale=1 / 8)
    ob.step(max_grad_norm=1.0)
    assert torch.allclose(a.param, b.param, rtol=1e-6, atol=1e-7)
    assert abs(oa.grad_norm(1 / 8) - ob.grad_norm()) < 1e-5 * ob.grad_norm()


@pytest.mark.parametrize("path", ["fp32", "bf16_cuda_cores", "bf16_tensor_cores"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 9, 7, 96, 80, 3), (3, 8, 8, 64, 128, 1), (1, 16, 16, 128, 64, 3), (2, 5, 6, 20, 12, 3),
                                              (2, 32, 32, 320, 320, 3), (4, 16, 16, 640, 328, 1), (1, 3, 3, 8, 8, 3)])
def test_conv_wgrad_vs_autograd(ops, path, B, H, W, Cin, Cout, k):
    """Weight + bias gradient against float64 autograd of F.conv2d on the same (already rounded) operands.  The tensor-core
    path (split-K mma.sync, several K slices at the larger shapes) and the CUDA-core path must agree with it alike."""
    if path == "bf16_tensor_cores" and (Cin % 8 or Cout % 8):
        pytest.skip("ragged channel counts take the CUDA-core kernel")
    dtype = torch.float32 if path == "fp32" else bf16
    x = _randn(B, H, W, Cin, seed=1, dtype=dtype)
    dy = _randn(B, H, W, Cout, seed=2, dtype=dtype)
    w = _randn(Cout, Cin, k, k, seed=3)
    _, dw_ref, db_ref = T.conv_grads(x.float().permute(0, 3, 1, 2).cpu().numpy(), w.cpu().numpy(),
                                     dy.float().permute(0, 3, 1, 2).cpu().numpy())
    dw_ref = torch.from_numpy(dw_ref.transpose(0, 2, 3, 1).reshape(Cout, -1))       # packed K order (kh, kw, ci)
    dw = torch.full((Cout, k * k * Cin), float("nan"), device="cuda")
    db = torch.full((Cout,), float("nan"), device="cuda")
    kw = dict(B=B, H=H, W=W, ksize=k, cuda_cores=(path == "bf16_cuda_cores"))
    ops.conv_wgrad(x, dy, dw, db, **kw)
    # products of the rounded operands are exact in fp32; what differs is the fp32 accumulation order over B*H*W pixels
    tol = 2e-5 if path == "bf16_tensor_cores" else 5e-6
    assert rel(dw.cpu(), dw_ref) < tol
    assert rel(db.cpu(), torch.from_numpy(db_ref)) < tol
    first = dw.clone()
    ops.conv_wgrad(x, dy, dw, db, accumulate=True, **kw)
    assert rel(dw.cpu(), 2 * dw_ref) < tol and rel(db.cpu(), 2 * torch.from_numpy(db_ref)) < tol
    ops.conv_wgrad(x, dy, dw, db, **kw)
    assert torch.equal(dw, first)                          # deterministic (fixed-order split-K reduction)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 16, 16, 64, 128, 3), (2, 12, 12, 320, 64, 3), (1, 24, 24, 160, 64, 1)])
def test_conv_dgrad_runs_on_the_igemm_kernel(ops, mode, B, H, W, Cin, Cout, k):
    """The data gradient of a stride-1 conv is the forward implicit GEMM over dy with the flipped / transposed weight."""
    dt = torch.float32 if mode == "fp32" else bf16
    w = _randn(Cout, Cin, k, k, seed=3, scale=(k * k * Cout) ** -0.5).to(dt).float()
    dy = _randn(B, H, W, Cout, seed=2, dtype=dt)
    x = np.zeros((B, Cin, H, W), np.float32)
    dx_ref, _, _ = T.conv_grads(x, w.cpu().numpy(), dy.float().permute(0, 3, 1, 2).cpu().numpy())
    with ops.precision(mode):
        wp = ops.pack_conv_dgrad_weight(w)
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=dt)
    ops.ConvPlan(dy, wp, dx, B=B, H=H, W=W, Cin=Cout, Cout=Cin, ksize=k).run()
    err = rel(dx.float().permute(0, 3, 1, 2).cpu(), torch.from_numpy(dx_ref))
    assert err < (1e-5 if mode == "fp32" else 4e-3)       # bf16: output rounding alone is ~1.5e-3


def test_one_layer_training_loop_end_to_end(ops):
    """conv3x3 'model', four optimizer steps entirely on the kernels (fp32 parity mode): forward plan -> MSE loss + gradient
    -> weight / bias gradient -> clip + AdamW; against F.conv2d + autograd + clip_grad_norm_ + torch.optim.AdamW."""
    from mirrorfusion_b200.train import B200AdamW, FlatParams, TrainLoss
    B, H, W, Cin, Cout = 2, 8, 8, 64, 64
    w0 = _randn(Cout, Cin, 3, 3, seed=1, scale=(9 * Cin) ** -0.5)
    b0 = _randn(Cout, seed=2, scale=0.1)
    flat = FlatParams({"weight": (Cout, 9 * Cin), "bias": (Cout,)}, "cuda", with_bf16=False)
    with ops.precision("fp32"):
        flat.p("weight").copy_(ops.pack_conv_weight(w0))
    flat.p("bias").copy_(b0)
    opt = B200AdamW(flat, lr=1e-2)
    tw, tb = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(b0.clone())
    topt = torch.optim.AdamW([tw, tb], lr=1e-2)
    x = torch.empty(B, H, W, Cin, device="cuda")
    y = torch.empty(B, H, W, Cout, device="cuda")
    dy = torch.empty_like(y)
    plan = ops.ConvPlan(x, flat.p("weight"), y, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3, bias=flat.p("bias"))
    L = TrainLoss(B, "cuda")
    for step in range(4):
        x.copy_(_randn(B, H, W, Cin, seed=10 + step))
        target = _randn(B, H, W, Cout, seed=20 + step)
        plan.run()
        loss = L(y.view(B, -1), target.view(B, -1), grad=dy.view(B, -1))
        ops.conv_wgrad(x, dy, flat.g("weight"), flat.g("bias"), B=B, H=H, W=W, ksize=3)
        opt.step(max_grad_norm=1.0)
        ty = F.conv2d(x.permute(0, 3, 1, 2), tw, tb, padding=1)
        tloss = F.mse_loss(ty, target.permute(0, 3, 1, 2))
        topt.zero_grad()
        tloss.backward()
        torch.nn.utils.clip_grad_norm_([tw, tb], 1.0)
        topt.step()
        assert abs(loss.item() - tloss.item()) < 1e-5 * tloss.item(), step
        # Adam's update lr * m / (sqrt(v) + eps) is ~ lr * sign(g): smooth everywhere except for the rare elements with
        # |g| ~ eps = 1e-8, where a 1e-10 summation-order difference between the two weight-gradient kernels moves the update
        # by up to lr * 1e-2 (measured: 1.6e-5 rel-L2 after the first step, from less than one such element of 36 864)
        assert rel(flat.p("weight"), tw.detach().permute(0, 2, 3, 1).reshape(Cout, -1)) < 1e-4, step
        assert rel(flat.p("bias"), tb.detach()) < 1e-4, step


@pytest.mark.parametrize("dtype", [torch.float32, bf16])
@pytest.mark.parametrize("B,HW,C1,C2,groups,silu,eps", [(2, 64, 320, 0, 32, True, 1e-5), (3, 100, 64, 0, 32, False, 1e-6),
                                                       (2, 256, 640, 320, 32, True, 1e-5), (1, 64, 1280, 1280, 32, True, 1e-5),
                                                       (2, 36, 128, 0, 32, True, 1e-5)])
def test_groupnorm_silu_backward_vs_autograd(ops, dtype, B, HW, C1, C2, groups, silu, eps):
    """(2, 256, 640, 320): the 960-channel skip concat, group 21 straddles the two source tensors."""
    C = C1 + C2
    x1 = _randn(B, HW, C1, seed=1, dtype=dtype) * 1.5 + 0.3
    x2 = (_randn(B, HW, C2, seed=2, dtype=dtype) * 0.7 - 0.2) if C2 else None
    dy = _randn(B, HW, C, seed=3, dtype=dtype)
    gamma, beta = 1.0 + 0.2 * _randn(C, seed=4), 0.1 * _randn(C, seed=5)
    xcat = x1 if x2 is None else torch.cat([x1, x2], -1)
    to_nchw = lambda t: t.float().permute(0, 2, 1).reshape(B, -1, HW, 1).cpu().numpy()
    dx_ref, dg_ref, db_ref = T.groupnorm_silu_grads(to_nchw(xcat), gamma.cpu().numpy(), beta.cpu().numpy(), to_nchw(dy), groups, eps, silu)
    dx_ref = torch.from_numpy(dx_ref).reshape(B, C, HW).permute(0, 2, 1)
    dx1 = torch.full_like(x1, float("nan"))
    dx2 = torch.full_like(x2, float("nan")) if C2 else None
    dg, db = torch.full((C,), float("nan"), device="cuda"), torch.full((C,), float("nan"), device="cuda")
    ws = torch.zeros(2 * B * C, device="cuda")
    ops.groupnorm_bwd(x1, x2, dy, gamma, beta, dx1, dx2, ws, B=B, HW=HW, groups=groups, eps=eps, silu=silu, dgamma=dg, dbeta=db)
    dx = dx1 if dx2 is None else torch.cat([dx1, dx2], -1)
    tol = 2e-5 if dtype == torch.float32 else 4e-3       # bf16: rounding of the stored dx alone is ~1.5e-3
    assert rel(dx.float().cpu(), dx_ref) < tol
    assert rel(dg.cpu(), torch.from_numpy(dg_ref)) < 2e-5 and rel(db.cpu(), torch.from_numpy(db_ref)) < 2e-5
    again = torch.full_like(x1, float("nan"))
    ops.groupnorm_bwd(x1, x2, dy, gamma, beta, again, dx2, ws, B=B, HW=HW, groups=groups, eps=eps, silu=silu, dgamma=dg, dbeta=db,
                      accumulate=True)
    assert torch.equal(again, dx1)                         # deterministic
    assert rel(dg.cpu(), 2 * torch.from_numpy(dg_ref)) < 2e-5


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 32, 32, 320, 320), (3, 8, 12, 128, 64)])
def test_downsample_dgrad_runs_on_the_up2x_plan(ops, mode, B, H, W, Cin, Cout):
    """Data gradient of Downsample2D's stride-2 conv3x3 (S/models/downsampling.py:146-152) = the sub-pixel up2x plan over dy
    with the parity-selected taps (ops.pack_conv_s2_dgrad_weight); H, W are the forward INPUT size."""
    dt = torch.float32 if mode == "fp32" else bf16
    w = _randn(Cout, Cin, 3, 3, seed=3, scale=(9 * Cout) ** -0.5).to(dt).float()
    dy = _randn(B, H // 2, W // 2, Cout, seed=2, dtype=dt)
    x = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double().cpu(), stride=2, padding=1).backward(dy.double().permute(0, 3, 1, 2).cpu())
    with ops.precision(mode):
        wp = ops.pack_conv_s2_dgrad_weight(w)
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=dt)
    plan = ops.ConvPlan(dy, wp, dx, B=B, H=H // 2, W=W // 2, Cin=Cout, Cout=Cin, ksize=3, up2x=True)
    assert plan.launches == 4
    plan.run()
    err = rel(dx.float().permute(0, 3, 1, 2).cpu(), x.grad)
    assert err < (1e-5 if mode == "fp32" else 4e-3)


def test_rowsum_per_image(ops):
    for dt in (torch.float32, bf16):
        dy = _randn(3, 50, 72, seed=1, dtype=dt)
        out = torch.full((3, 72), float("nan"), device="cuda")
        ops.rowsum_per_image(dy, out, B=3, HW=50)
        assert rel(out, dy.float().sum(1)) < 1e-6


def test_groupnorm_backward_residual_addend(ops):
    B, HW, C = 2, 64, 128
    x, dy, dres = _randn(B, HW, C, seed=1), _randn(B, HW, C, seed=2), _randn(B, HW, C, seed=3)
    gamma, beta = 1.0 + 0.2 * _randn(C, seed=4), 0.1 * _randn(C, seed=5)
    ws = torch.zeros(2 * B * C, device="cuda")
    a, b = torch.empty_like(x), torch.empty_like(x)
    ops.groupnorm_bwd(x, None, dy, gamma, beta, a, None, ws, B=B, HW=HW, groups=32, eps=1e-5, silu=True)
    ops.groupnorm_bwd(x, None, dy, gamma, beta, b, None, ws, B=B, HW=HW, groups=32, eps=1e-5, silu=True, dres=dres)
    assert torch.allclose(b, a + dres, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["id", "sc"])
def test_resnet_block_forward_backward_vs_reference_autograd(ops, precision, tag):
    """One ResnetBlock2D trained on the kernels (mirrorfusion_b200/backward.py): forward, then backward twice (gradient
    accumulation) against float64 autograd of the oracle block, which tests/test_oracle_train.py pins to the reference's own
    ResnetBlock2D autograd.  fp32 parity mode: 5e-5; bf16 (two convs + two norms deep, bf16 activations and gradients): 3e-2."""
    from mirrorfusion_b200.backward import ResnetBlockTrainer, pack_resnet_state_dict, resnet_param_shapes, unpack_conv_grad
    from mirrorfusion_b200.train import FlatParams
    if precision == "fp32":     # the reference-golden geometries
        cin, cout = (64, 64) if tag == "id" else (64, 128)
    else:                       # the bf16 GroupNorm kernels take 4 or >= 8 channels per group (every SD1.5 / VAE level)
        cin, cout = (256, 256) if tag == "id" else (128, 256)
    sd, x, emb, d_out = T.resnet_block_case(cin, cout)
    ref = T.resnet_block_grads(sd, "r", x, emb, d_out)
    B, _, H, W = x.shape
    flat = FlatParams(resnet_param_shapes("r", cin, cout), "cuda")
    flat.load_state_dict(pack_resnet_state_dict("r", sd))
    blk = ResnetBlockTrainer(flat, "r", B=B, H=H, W=W, Cin=cin, Cout=cout, precision=precision)
    dt = torch.float32 if precision == "fp32" else bf16
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(B, H * W, -1).contiguous()
    tol = 5e-5 if precision == "fp32" else 3e-2
    out = blk.forward(nhwc(x).cuda().to(dt), ref["rowbias"].float().cuda())
    assert rel(out.float().cpu(), nhwc(ref["out"])) < (tol if precision == "fp32" else 1e-2)
    for rep in range(2):
        dx, drb = blk.backward(nhwc(d_out).cuda().to(dt))
    assert rel(dx.float().cpu(), nhwc(ref["dx"])) < tol
    assert rel(drb.cpu(), ref["d_rowbias"]) < tol
    for name in flat.table:
        want = ref[name]
        got = flat.g(name).cpu()
        got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
        assert rel(got, 2 * want) < tol, name


@pytest.mark.parametrize("path", ["fp32", "bf16_cuda_cores", "bf16_tensor_cores"])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 32, 32, 320, 320), (3, 8, 12, 24, 40)])
def test_conv_wgrad_stride2_vs_autograd(ops, path, B, H, W, Cin, Cout):
    """Weight / bias gradient of Downsample2D's stride-2 conv3x3 (x [B,H,W,Cin], dy [B,H/2,W/2,Cout])."""
    dtype = torch.float32 if path == "fp32" else bf16
    x = _randn(B, H, W, Cin, seed=1, dtype=dtype)
    dy = _randn(B, H // 2, W // 2, Cout, seed=2, dtype=dtype)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(Cout, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double().permute(0, 3, 1, 2).cpu(), w, b, stride=2, padding=1).backward(dy.double().permute(0, 3, 1, 2).cpu())
    dw_ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, -1)
    dw = torch.full((Cout, 9 * Cin), float("nan"), device="cuda")
    db = torch.full((Cout,), float("nan"), device="cuda")
    ops.conv_wgrad(x, dy, dw, db, B=B, H=H, W=W, ksize=3, stride=2, cuda_cores=(path == "bf16_cuda_cores"))
    tol = 2e-5 if path == "bf16_tensor_cores" else 5e-6
    assert rel(dw.cpu(), dw_ref) < tol and rel(db.cpu(), b.grad) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_downsample_forward_backward_vs_autograd(ops, precision):
    """Downsample2D trained on the kernels (backward.DownsampleTrainer) against float64 autograd of the oracle's downsample."""
    from mirrorfusion_b200.backward import DownsampleTrainer, unpack_conv_grad
    from mirrorfusion_b200.train import FlatParams
    from oracle import mf_oracle as O
    gen = torch.Generator().manual_seed(21)
    B, C, H, W = 2, 128, 16, 16
    sd = {"d.conv.weight": (torch.randn(C, C, 3, 3, generator=gen, dtype=torch.float64) * (9 * C) ** -0.5).requires_grad_(True),
          "d.conv.bias": (torch.randn(C, generator=gen, dtype=torch.float64) * 0.1).requires_grad_(True)}
    x = torch.randn(B, C, H, W, generator=gen, dtype=torch.float64, requires_grad=True)
    d_out = torch.randn(B, C, H // 2, W // 2, generator=gen, dtype=torch.float64)
    y = O.downsample(sd, "d", x)
    y.backward(d_out)
    flat = FlatParams({"d.conv.weight": (C, 9 * C), "d.conv.bias": (C,)}, "cuda")
    flat.load_state_dict({"d.conv.weight": sd["d.conv.weight"].detach().permute(0, 2, 3, 1).reshape(C, -1).float(),
                          "d.conv.bias": sd["d.conv.bias"].detach().float()})
    blk = DownsampleTrainer(flat, "d", B=B, H=H, W=W, C=C, precision=precision)
    dt = torch.float32 if precision == "fp32" else bf16
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).reshape(B, -1, C).contiguous()
    tol = 2e-5 if precision == "fp32" else 1e-2
    assert rel(blk.forward(nhwc(x).float().cuda().to(dt)).float().cpu(), nhwc(y)) < tol
    dx = blk.backward(nhwc(d_out).float().cuda().to(dt))
    assert rel(dx.float().cpu(), nhwc(x.grad)) < tol
    assert rel(unpack_conv_grad(flat.g("d.conv.weight"), 3).cpu(), sd["d.conv.weight"].grad) < tol
    assert rel(flat.g("d.conv.bias").cpu(), sd["d.conv.bias"].grad) < tol
