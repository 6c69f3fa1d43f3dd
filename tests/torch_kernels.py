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
                 res2=None, up2x=False, extras=()):
        kext = sum(e.shape[-1] for e in extras)
        assert tuple(w.shape) == ((4, Cout, 4 * Cin + kext) if up2x else (Cout, ksize * ksize * Cin + kext))
        self.a = (x, w, out, B, H, W, Cin, Cout, ksize, bias, rowbias, res1)
        self.res2 = res2
        self.extras = list(extras)          # extra 1x1 K-segments over tensors at OUTPUT resolution (shortcut halves)
        self.stride, self.up2x = stride, up2x
        self.launches = 4 if up2x else 1
        Ho, Wo = (2 * H, 2 * W) if up2x else (H // stride, W // stride)
        self.flops = 2.0 * B * Ho * Wo * Cout * (ksize * ksize * Cin + kext)

    def run(self):
        x, w, out, B, H, W, Cin, Cout, k, bias, rowbias, res1 = self.a
        if self.up2x:
            xp = F.pad(_nchw(x, B, H, W).float(), (1, 1, 1, 1))
            y = torch.zeros(B, Cout, 2 * H, 2 * W)
            for py in range(2):
                for px in range(2):
                    wk = w[py * 2 + px][:, :4 * Cin].reshape(Cout, 2, 2, Cin).permute(0, 3, 1, 2).float()
                    y[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], wk, bias)
            y = _nhwc(y).reshape(out.shape)
            if res1 is not None:
                y = y + res1.reshape(out.shape)
            if self.res2 is not None:
                y = y + self.res2.reshape(out.shape)
            out.copy_(y)
            return
        wk = w[:, :k * k * Cin].reshape(Cout, k, k, Cin).permute(0, 3, 1, 2)
        y = F.conv2d(_nchw(x, B, H, W).float(), wk.float(), bias, stride=self.stride, padding=k // 2)
        if rowbias is not None:
            y = y + rowbias[:, :Cout, None, None]
        y = _nhwc(y).reshape(out.shape)
        off = k * k * Cin
        for e in self.extras:
            ce = e.shape[-1]
            y = y + (e.reshape(-1, ce).float() @ w[:, off:off + ce].float().t()).reshape(out.shape)
            off += ce
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
                  dres=None, dres2=None, stats=None):
    x = (x1 if x2 is None else torch.cat([x1, x2], -1)).reshape(B, HW, -1).permute(0, 2, 1).detach().clone().requires_grad_(True)
    g, b = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    y = F.group_norm(x, groups, g, b, eps)
    if silu:
        y = F.silu(y)
    y.backward(dy.reshape(B, HW, -1).permute(0, 2, 1))
    dx = x.grad.permute(0, 2, 1)
    for extra in (dres, dres2):
        if extra is not None:
            dx = dx + extra.reshape(B, HW, -1)
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
    """NCHW fp32 inputs, w [3,3,Cin,Cout] -> NHWC out (and out_post = out + tap: the UNet's conv_in tap site)."""
    x = sample if cond is None else torch.cat([sample, cond], 1)
    y = F.conv2d(x.float(), w.permute(3, 2, 0, 1).float(), bias, padding=1)
    out.copy_(_nhwc(y).reshape(out.shape))
    if out_post is not None:
        out_post.copy_(out + tap.reshape(out.shape))


def conv_out(x, w, bias, out, *, B, H, W):
    """x NHWC, w [Cout, 3, 3, Cin] -> out NCHW fp32."""
    out.copy_(F.conv2d(_nchw(x, B, H, W).float(), w.permute(0, 3, 1, 2).float(), bias, padding=1))


def conv_out_bwd(dy, w, dx, *, B, H, W):
    xi = torch.zeros(B, w.shape[-1], H, W, requires_grad=True)
    F.conv2d(xi, w.permute(0, 3, 1, 2).float(), None, padding=1).backward(dy)
    dx.copy_(_nhwc(xi.grad).reshape(dx.shape))


def dgrad_repack(wp, wd, ksize):
    Cout, Cin = wp.shape[0], wd.shape[0]
    wd.copy_(wp.view(Cout, ksize, ksize, Cin).flip(1, 2).permute(3, 1, 2, 0).reshape(Cin, -1).to(wd.dtype))


def sumpool2x2(du, dx, *, B, H, W):
    v = du.reshape(B, H, 2, W, 2, -1)
    dx.copy_(v.sum((2, 4)).reshape(dx.shape))


def layernorm(x, gamma, beta, out, eps=1e-5):
    out.copy_(F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps))


def layernorm_bwd(x, dy, gamma, dx, eps=1e-5, dres=None):
    xi = x.detach().float().clone().requires_grad_(True)
    F.layer_norm(xi, (x.shape[-1],), gamma, torch.zeros_like(gamma), eps).backward(dy.float())
    dx.copy_(xi.grad + (0 if dres is None else dres))


def geglu(proj, out=None, d_out=None, d_proj=None):
    pi = proj.detach().float().clone().requires_grad_(True)
    h, gate = pi.chunk(2, -1)
    y = h * F.gelu(gate)
    if out is not None:
        out.copy_(y.detach())
    if d_proj is not None:
        y.backward(d_out.float())
        d_proj.copy_(pi.grad)


def _heads(t, ld, B, T, heads, d):
    """[B, T, ld]-strided view (possibly a flat offset view into a fused buffer) -> [B, heads, T, d]."""
    flat = t.reshape(-1)
    return torch.as_strided(flat, (B, T, heads, d), (T * ld, ld, d, 1)).permute(0, 2, 1, 3)


_LOG2E = 1.4426950408889634


def attention_lse(q, k, v, out, lse, *, B, heads, head_dim, Tq, Tk, ldq=None, ldk=None, ldv=None, ldo=None):
    d = head_dim
    qh, kh, vh = _heads(q, ldq, B, Tq, heads, d).float(), _heads(k, ldk, B, Tk, heads, d).float(), _heads(v, ldv, B, Tk, heads, d).float()
    s = (qh @ kh.transpose(-1, -2)) * d ** -0.5
    o = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, Tq, heads * d)
    torch.as_strided(out.reshape(-1), (B, Tq, heads * d), (Tq * ldo, ldo, 1)).copy_(o)
    lse.copy_((torch.logsumexp(s, -1) * _LOG2E).reshape(-1))


def attention_bwd(q, k, v, o, d_o, lse, dvec, dq, dk=None, dv=None, *, B, heads, head_dim, Tq, Tk, ldq=None, ldk=None, ldv=None, ldo=None,
                  lddo=None, lddq=None, lddk=None, lddv=None):
    d = head_dim
    qh = _heads(q, ldq, B, Tq, heads, d).detach().float().clone().requires_grad_(True)
    kh = _heads(k, ldk, B, Tk, heads, d).detach().float().clone().requires_grad_(True)
    vh = _heads(v, ldv, B, Tk, heads, d).detach().float().clone().requires_grad_(True)
    s = (qh @ kh.transpose(-1, -2)) * d ** -0.5
    (torch.softmax(s, -1) @ vh).backward(_heads(d_o, lddo, B, Tq, heads, d).float())
    for dst, g, ld, T in ((dq, qh.grad, lddq, Tq), (dk, kh.grad, lddk, Tk), (dv, vh.grad, lddv, Tk)):
        if dst is not None:
            torch.as_strided(dst.reshape(-1), (B, T, heads, d), (T * ld, ld, d, 1)).copy_(g.permute(0, 2, 1, 3))


def nchw_to_nhwc(x, out):
    out.copy_(_nhwc(x).reshape(out.shape))
