"""CPU stand-in for the kernel namespace `mirrorfusion_b200.ops` — TEST INFRASTRUCTURE ONLY.  Same call signatures, plain torch fp32
math on NHWC tensors, so that launch PROGRAMS written against the kernels (mirrorfusion_b200/backward.py) can be checked for their
dataflow against the reference's autograd without a GPU.  It is never imported by the package; the product path has no CPU
fallback (ops raises without CUDA)."""
import torch
import torch.nn.functional as F

from mirrorfusion_b200.ops import gn_ws_floats, pack_conv_weight  # noqa: F401  (pure host functions)


def _nchw(t, B, H, W):
    return t.reshape(B, H, W, -1).permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1)


class ConvPlan:
    """Subset of mfb_conv_desc: packed weight [Cout, k*k*Cin], stride 1 / 2, bias, rowbias, res1, and the `up2x` sub-pixel form
    ([4, Cout, 4*Cin]: phase (py, px) is a 2x2 conv over the low-resolution input at offsets {-1, 0} / {0, +1})."""

    def __init__(self, x, w, out, *, B, H, W, Cin, Cout, ksize=1, stride=1, bias=None, rowbias=None, rowbias_ld=0, res1=None,
                 res2=None, up2x=False):
        assert tuple(w.shape) == ((4, Cout, 4 * Cin) if up2x else (Cout, ksize * ksize * Cin))
        self.a = (x, w, out, B, H, W, Cin, Cout, ksize, bias, rowbias, res1)
        self.res2 = res2
        self.stride, self.up2x = stride, up2x
        self.launches = 4 if up2x else 1

    def run(self):
        x, w, out, B, H, W, Cin, Cout, k, bias, rowbias, res1 = self.a
        if self.up2x:
            xp = F.pad(_nchw(x, B, H, W).float(), (1, 1, 1, 1))
            y = torch.zeros(B, Cout, 2 * H, 2 * W)
            for py in range(2):
                for px in range(2):
                    wk = w[py * 2 + px].view(Cout, 2, 2, Cin).permute(0, 3, 1, 2).float()
                    y[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], wk, bias)
            out.copy_(_nhwc(y).reshape(out.shape))
            return
        wk = w.view(Cout, k, k, Cin).permute(0, 3, 1, 2)
        y = F.conv2d(_nchw(x, B, H, W).float(), wk.float(), bias, stride=self.stride, padding=k // 2)
        if rowbias is not None:
            y = y + rowbias[:, :Cout, None, None]
        y = _nhwc(y).reshape(out.shape)
        if res1 is not None:
            y = y + res1.reshape(out.shape)
        if self.res2 is not None:
            y = y + self.res2.reshape(out.shape)
        out.copy_(y)


def groupnorm(x1, x2, gamma, beta, out, stats_ws, *, B, HW, groups, eps, silu, part1=None, part2=None):
    x = x1 if x2 is None else torch.cat([x1, x2], -1)
    y = F.group_norm(x.reshape(B, HW, -1).permute(0, 2, 1), groups, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    out.copy_(y.permute(0, 2, 1).reshape(out.shape))


def groupnorm_bwd(x1, x2, dy, gamma, beta, dx1, dx2, ws, *, B, HW, groups, eps, silu, dgamma=None, dbeta=None, accumulate=False,
                  dres=None):
    x = (x1 if x2 is None else torch.cat([x1, x2], -1)).reshape(B, HW, -1).permute(0, 2, 1).detach().clone().requires_grad_(True)
    g, b = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    y = F.group_norm(x, groups, g, b, eps)
    if silu:
        y = F.silu(y)
    y.backward(dy.reshape(B, HW, -1).permute(0, 2, 1))
    dx = x.grad.permute(0, 2, 1)
    if dres is not None:
        dx = dx + dres.reshape(B, HW, -1)
    C1 = x1.shape[-1]
    dx1.copy_(dx[..., :C1].reshape(dx1.shape))
    if dx2 is not None:
        dx2.copy_(dx[..., C1:].reshape(dx2.shape))
    for dst, src in ((dgamma, g.grad), (dbeta, b.grad)):
        if dst is not None:
            dst.copy_(src + (dst if accumulate else 0))


def conv_wgrad(x, dy, dw, dbias=None, *, B, H, W, ksize, stride=1, accumulate=False, ws=None, cuda_cores=False):
    Cin, Cout = x.shape[-1], dy.shape[-1]
    w = torch.zeros(Cout, Cin, ksize, ksize, requires_grad=True)
    b = torch.zeros(Cout, requires_grad=True)
    F.conv2d(_nchw(x, B, H, W).float(), w, b, stride=stride, padding=ksize // 2).backward(_nchw(dy, B, H // stride, W // stride).float())
    gw = w.grad.permute(0, 2, 3, 1).reshape(Cout, -1)
    dw.copy_(gw + (dw if accumulate else 0))
    if dbias is not None:
        dbias.copy_(b.grad + (dbias if accumulate else 0))


def rowsum_per_image(dy, out, *, B, HW):
    out.copy_(dy.reshape(B, HW, -1).float().sum(1))


def upsample2x(x, out, *, B, H, W):
    """nearest x2 of NHWC pixel vectors (dtype-agnostic copy, like the kernel)."""
    v = x.reshape(B, H, W, -1)
    out.copy_(v.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(out.shape))


def timestep_sinusoid(t, out):
    from oracle import mf_oracle as O
    out.copy_(O.timestep_embedding(t, out.shape[-1]))


def linear_small(x, w, b, y, act_in=False, act_out=False):
    v = F.silu(x) if act_in else x
    r = F.linear(v.float(), w.float(), b)
    y.copy_(F.silu(r) if act_out else r)


def silu_bwd(x, dy=None, y=None, dx=None):
    if y is not None:
        y.copy_(F.silu(x))
    if dx is not None:
        sg = torch.sigmoid(x)
        dx.copy_(dy.float() * (sg * (1 + x * (1 - sg))))


def f32_to_bf16(x, out):
    out.copy_(x.view_as(out))


def conv_in(sample, cond, w, bias, out, tap=None, out_post=None):
    """NCHW fp32 inputs, w [3,3,Cin,Cout] -> NHWC out."""
    x = sample if cond is None else torch.cat([sample, cond], 1)
    y = F.conv2d(x.float(), w.permute(3, 2, 0, 1).float(), bias, padding=1)
    out.copy_(_nhwc(y).reshape(out.shape))


def nchw_to_nhwc(x, out):
    out.copy_(_nhwc(x).reshape(out.shape))
