"""Torch-tensor front ends of the C-ABI kernels.  PyTorch only supplies device memory and the stream;
all math runs in libmfb200.so.  Every function requires CUDA tensors on an sm_100 device."""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ConvDesc, check

bf16 = torch.bfloat16
f32 = torch.float32

# Storage dtype of activations and packed weights.  bf16 = the product path (tcgen05 kernels).  float32 = the fp32
# PARITY MODE (BASELINE config 1, rel-L2 1e-4 bar): the same host program on the CUDA-core kernels of csrc/fp32mode.cu.
# Engines read it while they are BUILT (buffers, packed weights); at run time every op dispatches on its tensors' dtype.
_ACT = [bf16]


def act_dtype():
    return _ACT[0]


@contextlib.contextmanager
def precision(mode: str):
    """`with ops.precision("fp32"): eng = StepEngine(...)` builds an engine in fp32 parity mode ("bf16" = default)."""
    if mode not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {mode!r}")
    old = _ACT[0]
    _ACT[0] = f32 if mode == "fp32" else bf16
    try:
        yield
    finally:
        _ACT[0] = old


def _is32(t: torch.Tensor) -> bool:
    if t.dtype not in (bf16, f32):
        raise ValueError(f"activation tensors must be bfloat16 (product path) or float32 (parity mode), got {t.dtype}")
    return t.dtype == f32


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def lib():
    if not torch.cuda.is_available():
        raise _lib.MfbError("mirrorfusion_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return _lib.load(torch.cuda.current_device())


def _req(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous CUDA {dtype} tensor, got {t.dtype} {t.device} contiguous={t.is_contiguous()}")


# --------------------------------------------------------------------------------------------- recorded launch programs
class Program:
    """`with ops.Program() as prog: <launches>` executes the launches and records them inside libmfb200 (mfb_program_begin / _end);
    `prog.run()` replays the whole sequence from one C call — no Python between the kernels."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p()

    def __enter__(self):
        check(self._L.mfb_program_begin(C.byref(self._h)))
        return self

    def __exit__(self, *exc):
        check(self._L.mfb_program_end())
        return False

    def __len__(self):
        return int(self._L.mfb_program_size(self._h))

    def run(self):
        check(self._L.mfb_program_run(self._h, _stream()))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.mfb_program_destroy(self._h)
                self._h = None
        except Exception:
            pass


def copy_f32(dst, src):
    """dst[:] = src (fp32, contiguous) by a library kernel, so the copy is part of a recorded program."""
    _req(dst, f32, "dst"); _req(src, f32, "src")
    if dst.numel() != src.numel():
        raise ValueError("copy_f32: size mismatch")
    check(lib().mfb_copy_f32(_ptr(dst), _ptr(src), dst.numel(), _stream()))


# --------------------------------------------------------------------------------------------- weight packing
def pack_conv_weight(w: torch.Tensor, extras: Sequence[torch.Tensor] = ()) -> torch.Tensor:
    """OIHW conv weight (or [out,in] linear weight) -> [Cout, kh*kw*Cin (+ extra 1x1 segments)] bf16,
    K ordered (kh, kw, cin) to match the tap-major K loop of the implicit GEMM."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    parts = [w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)]
    for e in extras:
        parts.append(e.reshape(e.shape[0], -1))
    return torch.cat(parts, 1).to(act_dtype()).contiguous()


def pack_upconv_weight(w: torch.Tensor, extras: Sequence[torch.Tensor] = ()) -> torch.Tensor:
    """3x3 conv applied to a nearest-2x upsampled input == four sub-pixel phases, each a 2x2 conv over the
    low-resolution input whose taps are sums of the 3x3 taps that land on the same source pixel.
    OIHW [Cout,Cin,3,3] -> [4 (py*2+px), Cout, 2*2*Cin (+ extra 1x1 segments)] bf16 (sums in fp32, one rounding)."""
    w = w.float()
    # rows: phase 0 -> {dy=-1: kh0, dy=0: kh1+kh2}; phase 1 -> {dy=0: kh0+kh1, dy=+1: kh2}   (same for columns)
    comb = [[(0,), (1, 2)], [(0, 1), (2,)]]
    out = []
    for py in range(2):
        for px in range(2):
            taps = []
            for ty in range(2):
                for tx in range(2):
                    acc = 0
                    for kh in comb[py][ty]:
                        for kw in comb[px][tx]:
                            acc = acc + w[:, :, kh, kw]
                    taps.append(acc)                      # [Cout, Cin]
            parts = [torch.stack(taps, 1).reshape(w.shape[0], -1)] + [e.reshape(e.shape[0], -1).float() for e in extras]
            out.append(torch.cat(parts, 1))
    return torch.stack(out, 0).to(act_dtype()).contiguous()


def pack_geglu(w: torch.Tensor, b: torch.Tensor):
    """GEGLU proj [8C, C]: rows [0,4C) value, [4C,8C) gate (activations.py:100-103) -> interleave per 128:
    64 value rows then the 64 matching gate rows."""
    half = w.shape[0] // 2
    wv, wg = w[:half].reshape(half // 64, 64, -1), w[half:].reshape(half // 64, 64, -1)
    wp = torch.stack([wv, wg], 1).reshape(2 * half, -1)
    bp = torch.stack([b[:half].reshape(-1, 64), b[half:].reshape(-1, 64)], 1).reshape(-1)
    return wp.to(act_dtype()).contiguous(), bp.float().contiguous()


# --------------------------------------------------------------------------------------------- implicit GEMM
class ConvPlan:
    """A prepared implicit-GEMM launch (tensor maps encoded once; all buffers at fixed addresses)."""

    def __init__(self, x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, B: int, H: int, W: int, Cin: int,
                 Cout: int, ksize: int = 1, stride: int = 1, extras: Sequence[torch.Tensor] = (),
                 bias: Optional[torch.Tensor] = None, rowbias: Optional[torch.Tensor] = None, rowbias_ld: int = 0,
                 alpha: Optional[torch.Tensor] = None, res1: Optional[torch.Tensor] = None,
                 res2: Optional[torch.Tensor] = None, geglu: bool = False, block_n: int = 0, igemm_mode: int = 0, up2x: bool = False,
                 pad0: bool = False):
        L = lib()
        dt = f32 if _is32(x) else bf16      # fp32 tensors select the parity-mode plan (mfb_conv_desc.dtype = 1)
        _req(x, dt, "x"); _req(w, dt, "w"); _req(out, dt, "out")
        d = ConvDesc()
        d.dtype = 1 if dt == f32 else 0
        d.B, d.H, d.W, d.Cin, d.Cout, d.ksize, d.stride = B, H, W, Cin, Cout, ksize, stride
        d.x, d.w, d.out = x.data_ptr(), w.data_ptr(), out.data_ptr()
        d.n_extra = len(extras)
        for i, e in enumerate(extras):
            _req(e, dt, f"extra{i}")
            d.extra_x[i] = e.data_ptr()
            d.extra_C[i] = e.shape[-1]
        for name, t in (("bias", bias), ("rowbias", rowbias), ("alpha", alpha)):
            if t is not None:
                if name == "rowbias":   # may be a column slice of a wider [B, ld] table
                    if t.dtype != torch.float32 or not t.is_cuda or t.stride(-1) != 1:
                        raise ValueError("rowbias: expected CUDA float32 with unit inner stride")
                else:
                    _req(t, torch.float32, name)
                setattr(d, name, t.data_ptr())
        d.rowbias_ld = rowbias_ld
        for name, t in (("res1", res1), ("res2", res2)):
            if t is not None:
                _req(t, dt, name)
                setattr(d, name, t.data_ptr())
        d.geglu = int(geglu)
        d.block_n = block_n
        d.igemm_mode = igemm_mode
        d.up2x = int(up2x)
        d.pad0 = int(pad0)
        ktot = (4 if up2x else ksize * ksize) * Cin + sum(e.shape[-1] for e in extras)
        want = (4, Cout, ktot) if up2x else (Cout, ktot)
        if tuple(w.shape) != want:
            raise ValueError(f"packed weight shape {tuple(w.shape)} != {want}")
        h = C.c_void_p()
        check(L.mfb_conv_plan_create(C.byref(d), C.byref(h)))
        self._h = h
        self._L = L
        self._keep = (x, w, out, extras, bias, rowbias, alpha, res1, res2)
        self.flops = L.mfb_plan_flops(h)
        self.launches = L.mfb_plan_launches(h)
        self.mode = L.mfb_plan_igemm_mode(h)      # 0 independent CTAs, 1 / 2 pair modes, 3 split-K pair
        Ho, Wo = (2 * H, 2 * W) if up2x else ((H + stride - 1) // stride, (W + stride - 1) // stride)
        # MMA work actually issued: the sub-pixel upsample plan runs 4 taps per output pixel where `flops` counts the 9 of the
        # reference's conv over the upsampled tensor
        self.exec_flops = 2.0 * B * Ho * Wo * Cout * ktot
        self.note = (f"M={B * Ho * Wo} N={Cout} K={ktot} k{ksize}" + (" s2" if stride == 2 else "") + (" up2x" if up2x else "") +
                     (" geglu" if geglu else "") + (" +res1" if res1 is not None else "") + (" +res2" if res2 is not None else "") +
                     (f" +{len(extras)}seg" if extras else ""))

    def run(self):
        check(self._L.mfb_plan_run(self._h, _stream()))

    def enable_output_stats(self, buf: Optional[torch.Tensor] = None):
        """Let the epilogue also emit the GroupNorm statistics of the output (per image, tile, channel).  Returns
        (buffer, tiles) or None if this plan cannot (GEGLU / tiles straddling images)."""
        n = self._L.mfb_plan_stats_floats(self._h)
        if n <= 0:
            return None
        if buf is None:
            buf = torch.zeros(n, device=self._keep[2].device, dtype=torch.float32)
        elif buf.numel() < n or buf.dtype != torch.float32:
            raise ValueError("statistics buffer too small")
        check(self._L.mfb_plan_set_stats(self._h, C.c_void_p(buf.data_ptr())))
        self._stats = buf
        return buf, self._L.mfb_plan_stats_tiles(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.mfb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def linear_plan(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, **kw) -> ConvPlan:
    """x [M, K] bf16, w [N, K] packed bf16 -> out [M, N] (or [M, N/2] with geglu)."""
    M, K = x.shape
    return ConvPlan(x, w, out, B=1, H=1, W=M, Cin=K, Cout=w.shape[0], ksize=1, **kw)


# --------------------------------------------------------------------------------------------- norms
GN_MAX_CHUNKS = 64


def gn_ws_floats(B: int, groups: int) -> int:
    """MFB_GN_WS_FLOATS of include/mfb200.h: stats + per-chunk partials + ticket counters."""
    return 2 * B * groups * (1 + GN_MAX_CHUNKS) + B


def groupnorm(x1, x2, gamma, beta, out, stats_ws, *, B, HW, groups, eps, silu, part1=None, part2=None):
    """part1 / part2: (buffer, tiles) from ConvPlan.enable_output_stats of the GEMM that produced x1 / x2."""
    L = lib()
    C1 = x1.shape[-1]
    C2 = 0 if x2 is None else x2.shape[-1]
    if _is32(x1):
        check(L.mfb_groupnorm_f32(_ptr(x1), C1, _ptr(x2), C2, B, HW, groups, eps, _ptr(gamma), _ptr(beta), int(silu), _ptr(out),
                                  _stream()))
        return
    if part1 is not None and (x2 is None or part2 is not None):
        check(L.mfb_groupnorm_prestat(_ptr(x1), C1, _ptr(part1[0]), part1[1], _ptr(x2), C2,
                                      _ptr(part2[0]) if part2 else None, part2[1] if part2 else 0, B, HW, groups, eps,
                                      _ptr(gamma), _ptr(beta), int(silu), _ptr(stats_ws), _ptr(out), _stream()))
        return
    check(L.mfb_groupnorm(_ptr(x1), C1, _ptr(x2), C2, B, HW, groups, eps, _ptr(gamma), _ptr(beta), int(silu),
                          _ptr(stats_ws), _ptr(out), _stream()))


def layernorm(x, gamma, beta, out, eps=1e-5):
    L = lib()
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    fn = L.mfb_layernorm_f32 if _is32(x) else L.mfb_layernorm
    check(fn(_ptr(x), rows, Cc, eps, _ptr(gamma), _ptr(beta), _ptr(out), _stream()))


def softmax_rows(x, out):
    """Row softmax of a [rows, cols] bf16 matrix (fp32 math)."""
    _req(x, bf16, "x"); _req(out, bf16, "out")
    check(lib().mfb_softmax_rows(_ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream()))


# --------------------------------------------------------------------------------------------- attention
_SCRATCH32 = {}


def _scratch32(key, n: int, device) -> torch.Tensor:
    """Grow-only fp32 scratch of the parity-mode backward kernels (one per role and device; launches on a stream serialise on it)."""
    t = _SCRATCH32.get(key)
    if t is None or t.numel() < n:
        t = _SCRATCH32[key] = torch.zeros(n, device=device, dtype=f32)
    return t


def add_f32(y, x):
    """y += x (fp32, parity mode)."""
    _req(y, f32, "y"); _req(x, f32, "x")
    check(lib().mfb_add_f32(_ptr(y), _ptr(x), y.numel(), _stream()))


def attention(q, k, v, out, *, B, heads, head_dim, Tq, Tk, ldq=None, ldk=None, ldv=None, ldo=None):
    """q [B,Tq,ldq], k / v [B,Tk,ld] (views into a fused q|k|v buffer are fine: pass the leading dimension)."""
    L = lib()
    fn = L.mfb_attention_f32 if _is32(q) else L.mfb_attention
    check(fn(_ptr(q), ldq or q.shape[-1], _ptr(k), ldk or k.shape[-1], _ptr(v), ldv or v.shape[-1],
             _ptr(out), ldo or out.shape[-1], B, heads, head_dim, Tq, Tk, _stream()))


def attention_lse(q, k, v, out, lse, *, B, heads, head_dim, Tq, Tk, ldq=None, ldk=None, ldv=None, ldo=None):
    """attention() that also keeps lse [B, heads, Tq] fp32 (log2-domain log-sum-exp per query row) for attention_bwd."""
    _req(lse, f32, "lse")
    if _is32(q):        # parity mode: the fp32 backward recomputes its own row statistics (attention_bwd below), lse stays unused
        attention(q, k, v, out, B=B, heads=heads, head_dim=head_dim, Tq=Tq, Tk=Tk, ldq=ldq, ldk=ldk, ldv=ldv, ldo=ldo)
        return
    check(lib().mfb_attention_lse(_ptr(q), ldq or q.shape[-1], _ptr(k), ldk or k.shape[-1], _ptr(v), ldv or v.shape[-1],
                                  _ptr(out), ldo or out.shape[-1], B, heads, head_dim, Tq, Tk, _ptr(lse), _stream()))


def attention_bwd(q, k, v, o, d_o, lse, dvec, dq, dk=None, dv=None, *, B, heads, head_dim, Tq, Tk, ldq=None, ldk=None, ldv=None, ldo=None,
                  lddo=None, lddq=None, lddk=None, lddv=None):
    """Backward of attention() on tcgen05 (bf16): dq always; dk / dv both or neither (cross attention to a frozen context).
    lse from attention_lse; dvec [B, heads, Tq] fp32 scratch."""
    if _is32(q):
        # fp32 parity mode (CUDA-core kernels, two deterministic passes that recompute the row statistics): o / lse / dvec unused;
        # a dq-only call (cross attention to the frozen context) gets throw-away dk / dv buffers
        ws = _scratch32(("attn_bwd_ws", q.device), 2 * B * heads * Tq, q.device)
        if dk is None:
            dk = _scratch32(("attn_bwd_dk", q.device), B * Tk * heads * head_dim, q.device)
            dv = _scratch32(("attn_bwd_dv", q.device), B * Tk * heads * head_dim, q.device)
            lddk = lddv = heads * head_dim
        check(lib().mfb_attention_bwd_f32(_ptr(q), ldq or q.shape[-1], _ptr(k), ldk or k.shape[-1], _ptr(v), ldv or v.shape[-1], _ptr(d_o),
                                          lddo or d_o.shape[-1], _ptr(dq), lddq or dq.shape[-1], _ptr(dk), lddk or dk.shape[-1], _ptr(dv),
                                          lddv or dv.shape[-1], _ptr(ws), B, heads, head_dim, Tq, Tk, _stream()))
        return
    for name, t in (("q", q), ("o", o), ("d_o", d_o), ("dq", dq)):
        if t.dtype != bf16 or not t.is_cuda:
            raise ValueError(f"{name}: expected a CUDA bfloat16 tensor")
    _req(lse, f32, "lse"); _req(dvec, f32, "dvec")
    if lse.numel() < B * heads * Tq or dvec.numel() < B * heads * Tq:
        raise ValueError("attention_bwd: lse / dvec smaller than B*heads*Tq")
    check(lib().mfb_attention_bwd(_ptr(q), ldq or q.shape[-1], _ptr(k), ldk or k.shape[-1], _ptr(v), ldv or v.shape[-1], _ptr(o),
                                  ldo or o.shape[-1], _ptr(d_o), lddo or d_o.shape[-1], _ptr(lse), _ptr(dvec), _ptr(dq),
                                  lddq or dq.shape[-1], _ptr(dk), 0 if dk is None else (lddk or dk.shape[-1]), _ptr(dv),
                                  0 if dv is None else (lddv or dv.shape[-1]), B, heads, head_dim, Tq, Tk, _stream()))


def transpose_tokens(x, out, *, ld, col0, Cc, B, T, ldt):
    check(lib().mfb_transpose_tokens(_ptr(x), ld, col0, Cc, B, T, _ptr(out), ldt, _stream()))


# --------------------------------------------------------------------------------------------- boundary / misc
def conv_in(sample, cond, w, bias, out, tap=None, out_post=None):
    B, Ca, H, W = sample.shape
    Cb = 0 if cond is None else cond.shape[1]
    fn = lib().mfb_conv_in_f32 if _is32(out) else lib().mfb_conv_in
    check(fn(_ptr(sample), Ca, _ptr(cond), Cb, B, H, W, _ptr(w), _ptr(bias), out.shape[-1], _ptr(out), _ptr(tap), _ptr(out_post),
             _stream()))


def conv_out(x, w, bias, out, *, B, H, W):
    fn = lib().mfb_conv_out_f32 if _is32(x) else lib().mfb_conv_out
    check(fn(_ptr(x), x.shape[-1], B, H, W, _ptr(w), _ptr(bias), out.shape[1], _ptr(out), _stream()))


def upsample2x(x, out, *, B, H, W):
    check(lib().mfb_upsample2x(_ptr(x), B, H, W, x.shape[-1], _ptr(out), _stream()))


def nchw_to_nhwc(x, out):
    B, Cc, H, W = x.shape
    check(lib().mfb_nchw_f32_to_nhwc_bf16(_ptr(x), B, Cc, H, W, _ptr(out), _stream()))


def nhwc_to_nchw(x, out):
    B, Cc, H, W = out.shape
    check(lib().mfb_nhwc_bf16_to_nchw_f32(_ptr(x), B, Cc, H, W, _ptr(out), _stream()))


def f32_to_bf16(x, out):
    """fp32 -> activation storage dtype (a plain copy in fp32 parity mode)."""
    if out.dtype == f32:
        out.copy_(x.view_as(out))
        return
    check(lib().mfb_f32_to_bf16(_ptr(x), x.numel(), _ptr(out), _stream()))


def timestep_sinusoid(t, out):
    check(lib().mfb_timestep_sinusoid(_ptr(t), t.numel(), out.shape[-1], _ptr(out), _stream()))


def linear_small(x, w, b, y, act_in=False, act_out=False):
    M, K = x.shape
    fn = lib().mfb_linear_small_f32 if _is32(w) else lib().mfb_linear_small
    check(fn(_ptr(x), M, K, _ptr(w), _ptr(b), w.shape[0], int(act_in), int(act_out), _ptr(y), _stream()))


def prep_image_u8(rgb_hwc, out_nchw):
    """uint8 [N,H,W,3] -> fp32 [N,3,H,W] in [-1,1]."""
    N, H, W, _ = rgb_hwc.shape
    check(lib().mfb_prep_image_u8(_ptr(rgb_hwc), N, H, W, _ptr(out_nchw), _stream()))


def prep_mask_depth(mask_u8, depth, mask_lat, depth_lat, scratch, factor=8, delta=0.5):
    """uint8 mask [N,H,W] (+ metric depth fp32 [N,H,W]) -> latent-resolution mask {0,1} / depth [-1,1] ([N,1,H/f,W/f] fp32)."""
    N, H, W = mask_u8.shape
    check(lib().mfb_prep_mask_depth(_ptr(mask_u8), _ptr(depth), N, H, W, factor, float(delta), _ptr(mask_lat), _ptr(depth_lat),
                                    _ptr(scratch), _stream()))


def resize_crop_bicubic(x, out, res: int, step: int = 1):
    """fp32 [..., Hs, Ws] -> [..., res/step, res/step]: torchvision Resize(res, BICUBIC, antialias) + CenterCrop(res), sampled every `step`."""
    Hs, Ws = x.shape[-2:]
    assert x.dtype == torch.float32 and x.is_contiguous() and tuple(out.shape[-2:]) == (res // step, res // step)
    check(lib().mfb_resize_crop_bicubic(_ptr(x), x.numel() // (Hs * Ws), Hs, Ws, res, step, _ptr(out), _stream()))


def depth_normalize(depth, mask_u8, out, scratch, delta=0.5):
    """metric depth fp32 [N,H,W] + uint8 mask [N,H,W] -> [-1,1] by the max depth over the mask (+ delta), same resolution."""
    N, H, W = depth.shape
    check(lib().mfb_depth_normalize(_ptr(depth), _ptr(mask_u8), N, H, W, float(delta), _ptr(out), _ptr(scratch), _stream()))


def post_image_u8(img_nchw, out_hwc):
    """fp32 [N,3,H,W] in [-1,1] -> uint8 [N,H,W,3]."""
    N, _, H, W = img_nchw.shape
    check(lib().mfb_post_image_u8(_ptr(img_nchw), N, H, W, _ptr(out_hwc), _stream()))


def latent_sample(mean, logvar, noise, scale, out):
    """out = scale * (mean + exp(0.5 * clamp(logvar)) * noise); noise=None -> the mode.  All fp32."""
    check(lib().mfb_latent_sample(_ptr(mean), _ptr(logvar), _ptr(noise), float(scale), _ptr(out), mean.numel(), _stream()))


def cfg_sched_step(eps_u, eps_c, x, last, m0, m1, coef):
    Bimg = x.shape[0]
    n = x.numel() // Bimg
    check(lib().mfb_cfg_sched_step(_ptr(eps_u), _ptr(eps_c), _ptr(x), _ptr(last), _ptr(m0), _ptr(m1), _ptr(coef), Bimg, n,
                                   _stream()))


# --------------------------------------------------------------------------------------------- training-step glue (config 4)
MSE_MAX_CHUNKS = 64          # MFB_MSE_MAX_CHUNKS
SQNORM_WS_FLOATS = 1184      # MFB_SQNORM_WS_FLOATS


def pack_conv_dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """Packed weight of the DATA gradient of a stride-1 conv / linear layer.  d x = conv(d y, W') with
    W'[ci, co, kh, kw] = W[co, ci, k-1-kh, k-1-kw] (a linear layer: W' = W^T), so the backward-data pass of every
    stride-1 conv3x3 / conv1x1 / linear is the SAME tcgen05 implicit GEMM run on d y: ConvPlan(dy, packed, dx,
    Cin=Cout_fwd, Cout=Cin_fwd).  OIHW or [out, in] in, [Cin_fwd, k*k*Cout_fwd] out."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    return pack_conv_weight(w.flip(2, 3).transpose(0, 1))


def pack_conv_s2_dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """Packed weight of the DATA gradient of the stride-2 conv3x3 (padding 1) of Downsample2D, for an `up2x` plan run on d y.
    Output pixel (2i+py, 2j+px) of d x only receives the taps whose parity matches: along one axis, py = 0 takes kh = 1 from
    d y[i]; py = 1 takes kh = 2 from d y[i] and kh = 0 from d y[i+1] — exactly the (phase, low-resolution offset) structure of
    the sub-pixel Upsample2D plan ({-1, 0} for phase 0, {0, +1} for phase 1), with the unused tap zero.  So Downsample2D's
    backward-data pass is ConvPlan(dy, packed, dx, B, H=Ho, W=Wo, Cin=Cout_fwd, Cout=Cin_fwd, ksize=3, up2x=True) on the
    tcgen05 kernel (9 of the 16 phase taps carry weight).  OIHW [Cout, Cin, 3, 3] -> [4, Cin, 4*Cout]."""
    w = w.float()
    k_of = [[None, 1], [2, 0]]          # [phase][tap] -> kernel index (tap 0 / 1 = the lower / upper low-res offset of that phase)
    zero = torch.zeros(w.shape[1], w.shape[0], dtype=w.dtype, device=w.device)
    out = []
    for py in range(2):
        for px in range(2):
            taps = []
            for ty in range(2):
                for tx in range(2):
                    kh, kw = k_of[py][ty], k_of[px][tx]
                    taps.append(zero if kh is None or kw is None else w[:, :, kh, kw].t())      # [Cin_fwd, Cout_fwd]
            out.append(torch.stack(taps, 1).reshape(w.shape[1], -1))
    return torch.stack(out, 0).to(act_dtype()).contiguous()


def add_noise(x0, noise, timesteps, alphas_cumprod, noisy=None, velocity=None):
    """DDPMScheduler.add_noise / get_velocity; x0 / noise [B, ...] fp32, timesteps [B] int64, all on the device."""
    _req(x0, f32, "x0"); _req(noise, f32, "noise"); _req(timesteps, torch.int64, "timesteps"); _req(alphas_cumprod, f32, "alphas_cumprod")
    B = x0.shape[0]
    check(lib().mfb_add_noise(_ptr(x0), _ptr(noise), _ptr(timesteps), _ptr(alphas_cumprod), alphas_cumprod.numel(), B,
                              x0.numel() // B, _ptr(noisy), _ptr(velocity), _stream()))


def mse_loss(pred, target, loss, ws, weights=None, per_sample=None, grad=None):
    """loss = mean_b(w_b mean_n (pred - target)^2) and (optionally) its gradient w.r.t. pred, one deterministic pass."""
    _req(pred, f32, "pred"); _req(target, f32, "target"); _req(loss, f32, "loss")
    B = pred.shape[0]
    if ws.numel() < B * MSE_MAX_CHUNKS:
        raise ValueError("mse_loss: workspace smaller than MFB_MSE_WS_FLOATS(B)")
    check(lib().mfb_mse_loss(_ptr(pred), _ptr(target), _ptr(weights), B, pred.numel() // B, _ptr(per_sample), _ptr(loss),
                             _ptr(grad), _ptr(ws), _stream()))


def grad_sqnorm(g, ws, out_sq, accumulate=False):
    _req(g, f32, "g")
    check(lib().mfb_grad_sqnorm(_ptr(g), g.numel(), _ptr(ws), _ptr(out_sq), int(accumulate), _stream()))


def adamw_step(param, grad, exp_avg, exp_avg_sq, hyper, param_bf16=None, grad_sqnorm=None, max_grad_norm=0.0):
    for name, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq), ("hyper", hyper)):
        _req(t, f32, name)
    check(lib().mfb_adamw_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(param_bf16), param.numel(),
                               _ptr(hyper), _ptr(grad_sqnorm), float(max_grad_norm), _stream()))


_WGRAD_WS: dict = {}


def conv_wgrad_ws_floats(B, H, W, Cin, Cout, ksize, stride=1) -> int:
    """Workspace (floats) of the tensor-core weight gradient for this layer geometry (H, W: the conv INPUT size)."""
    return int(_lib.load().mfb_conv_wgrad_tc_ws_floats(B, H, W, Cin, Cout, ksize, stride))


def conv_wgrad(x, dy, dw, dbias=None, *, B, H, W, ksize, stride=1, accumulate=False, ws=None, cuda_cores=False):
    """dw [Cout, k*k*Cin] fp32 (packed K order) and dbias [Cout] of a conv (stride 1, or 2 for the 3x3 of Downsample2D) / linear;
    x [B,H,W,Cin], dy [B,H/stride,W/stride,Cout] NHWC bf16 or fp32.
    bf16 operands with Cin, Cout multiples of 8 run on the tensor cores (ws: conv_wgrad_ws_floats(...) floats; if not given, a
    grow-only per-device workspace is used); fp32 operands (parity mode), ragged channel counts or cuda_cores=True take the CUDA-core kernel."""
    is32 = _is32(x)
    _req(dy, x.dtype, "dy"); _req(dw, f32, "dw")
    Cin, Cout = x.shape[-1], dy.shape[-1]
    if tuple(dw.shape) != (Cout, ksize * ksize * Cin):
        raise ValueError(f"dw shape {tuple(dw.shape)} != {(Cout, ksize * ksize * Cin)}")
    if not is32 and not cuda_cores and Cin % 8 == 0 and Cout % 8 == 0:
        need = conv_wgrad_ws_floats(B, H, W, Cin, Cout, ksize, stride)
        if ws is None:      # one grow-only workspace per device, shared by all layers (launches on one stream serialise on it)
            ws = _WGRAD_WS.get(x.device)
            if ws is None or ws.numel() < need:
                ws = _WGRAD_WS[x.device] = torch.empty(need, device=x.device, dtype=f32)
        _req(ws, f32, "ws")
        check(lib().mfb_conv_wgrad_tc(_ptr(x), _ptr(dy), B, H, W, Cin, Cout, ksize, stride, _ptr(dw), _ptr(dbias), int(accumulate),
                                      _ptr(ws), ws.numel(), _stream()))
        return
    check(lib().mfb_conv_wgrad(_ptr(x), _ptr(dy), int(is32), B, H, W, Cin, Cout, ksize, stride, _ptr(dw), _ptr(dbias), int(accumulate),
                               _stream()))


_GNB_WS: dict = {}


def groupnorm_stats(x1, x2, stats_ws, *, B, HW, groups):
    """stats_ws[0 : 2*B*groups] = per-(image, group) {sum, sum of squares} (bf16 inputs; workspace of gn_ws_floats(B, groups))."""
    _req(x1, bf16, "x1")
    check(lib().mfb_groupnorm_stats(_ptr(x1), x1.shape[-1], _ptr(x2), 0 if x2 is None else x2.shape[-1], B, HW, groups, _ptr(stats_ws),
                                    _stream()))


def groupnorm_bwd(x1, x2, dy, gamma, beta, dx1, dx2, ws, *, B, HW, groups, eps, silu, dgamma=None, dbeta=None, accumulate=False,
                  dres=None, dres2=None, stats=None):
    """Backward of groupnorm(): dx (per source tensor of the concat), dgamma, dbeta (None for a frozen layer).
    dres / dres2 [B, HW, C]: gradients of residual / skip paths, added to dx in the same pass.
    bf16 tensors with channel counts that are multiples of 8 take the vectorised two-pass kernels (mfb_groupnorm_bwd2; `stats`: the
    forward's statistics buffer if the caller kept it, otherwise they are recomputed; `ws` is ignored: a per-geometry workspace
    is kept); fp32 tensors (parity mode) the CUDA-core kernel (ws: 2*B*C floats, no dres2)."""
    is32 = _is32(x1)
    C1 = x1.shape[-1]
    C2 = 0 if x2 is None else x2.shape[-1]
    C = C1 + C2
    _req(dy, x1.dtype, "dy"); _req(dx1, x1.dtype, "dx1")
    if not is32 and C1 % 8 == 0 and C2 % 8 == 0 and (C // groups >= 8 or C // groups == 4):
        L = lib()
        # one workspace per (device, geometry): the ticket counters inside it sit at geometry-dependent offsets and must stay zero
        # between calls, so layouts never share memory (launches on one stream serialise on it); allocated at the first call
        key = (x1.device, B, C, groups)
        w2 = _GNB_WS.get(key)
        need = int(L.mfb_groupnorm_bwd2_ws_floats(B, C, groups))
        if w2 is None:
            w2 = _GNB_WS[key] = torch.zeros(need + gn_ws_floats(B, groups), device=x1.device, dtype=f32)
        st_ws = w2[need:]
        if stats is None:
            check(L.mfb_groupnorm_stats(_ptr(x1), C1, _ptr(x2), C2, B, HW, groups, _ptr(st_ws), _stream()))
            stats = st_ws
        check(L.mfb_groupnorm_bwd2(_ptr(x1), C1, _ptr(x2), C2, _ptr(dy), B, HW, groups, eps, _ptr(gamma), _ptr(beta), int(silu),
                                   _ptr(stats), _ptr(dres), _ptr(dres2), _ptr(dx1), _ptr(dx2), _ptr(dgamma), _ptr(dbeta), _ptr(w2),
                                   int(accumulate), _stream()))
        return
    if dres2 is not None and not is32:
        raise ValueError("groupnorm_bwd: dres2 needs the vectorised bf16 kernels or fp32 parity mode")
    if ws is None and is32:
        ws = _scratch32(("gn_bwd_ws", x1.device), 2 * B * C, x1.device)
    if ws.numel() < 2 * B * C:
        raise ValueError("groupnorm_bwd: workspace smaller than MFB_GN_BWD_WS_FLOATS(B, C)")
    check(lib().mfb_groupnorm_bwd(_ptr(x1), C1, _ptr(x2), C2, _ptr(dy), int(is32), B, HW, groups, eps, _ptr(gamma), _ptr(beta),
                                  int(silu), _ptr(dres), _ptr(dx1), _ptr(dx2), _ptr(dgamma), _ptr(dbeta), _ptr(ws), int(accumulate),
                                  _stream()))
    if dres2 is not None:       # parity mode: the skip-path gradient of the main input, added by its own launch
        add_f32(dx1, dres2)


def layernorm_bwd(x, dy, gamma, dx, eps=1e-5, dres=None):
    """Data gradient of layernorm() (+ dres, the gradient of the residual path), bf16 [rows, C]."""
    if _is32(x):
        layernorm_bwd_f32(x, dy, gamma, dx, eps)
        if dres is not None:
            add_f32(dx, dres)
        return
    for name, t in (("x", x), ("dy", dy), ("dx", dx)):
        _req(t, bf16, name)
    check(lib().mfb_layernorm_bwd(_ptr(x), _ptr(dy), x.numel() // x.shape[-1], x.shape[-1], eps, _ptr(gamma), _ptr(dres), _ptr(dx), _stream()))


def geglu(proj, out=None, d_out=None, d_proj=None):
    """Un-fused GEGLU on proj [rows, 2C] = [h | gate] (bf16): forward value and / or backward."""
    if _is32(proj):
        return geglu_f32(proj, out=out, d_out=d_out, d_proj=d_proj)
    _req(proj, bf16, "proj")
    Cc = proj.shape[-1] // 2
    check(lib().mfb_geglu(_ptr(proj), proj.numel() // proj.shape[-1], Cc, _ptr(out), _ptr(d_out), _ptr(d_proj), _stream()))


def conv_out_bwd(dy, w, dx, *, B, H, W):
    """Data gradient of conv_out: dy fp32 NCHW [B, Cout, H, W], w fp32 [Cout, 3, 3, Cin], dx bf16 [B, H*W, Cin]."""
    _req(dy, f32, "dy"); _req(w, f32, "w")
    if _is32(dx):
        _req(dx, f32, "dx")
        check(lib().mfb_conv_out_bwd_f32(_ptr(dy), B, H, W, dx.shape[-1], dy.shape[1], _ptr(w), _ptr(dx), _stream()))
        return
    _req(dx, bf16, "dx")
    check(lib().mfb_conv_out_bwd(_ptr(dy), B, H, W, dx.shape[-1], dy.shape[1], _ptr(w), _ptr(dx), _stream()))


def dgrad_repack(wp, wd, ksize):
    """wd [Cin, k*k*Cout] = data-gradient packing of the packed forward weight wp [Cout, k*k*Cin] (pack_conv_dgrad_weight of the
    unpacked weight), one kernel; bf16 on the device, anything else (fp32 parity mode) by torch."""
    Cout, Cin = wp.shape[0], wd.shape[0]
    if wp.dtype == bf16 and wd.dtype == bf16 and wp.is_cuda and wp.is_contiguous() and wd.is_contiguous():
        check(lib().mfb_dgrad_repack(_ptr(wp), Cout, Cin, ksize, _ptr(wd), _stream()))
        return
    w = wp.view(Cout, ksize, ksize, Cin).flip(1, 2).permute(3, 1, 2, 0)
    wd.copy_(w.reshape(Cin, -1).to(wd.dtype))


def sumpool2x2(du, dx, *, B, H, W):
    """dx [B, H*W, C] = 2x2 sums of du [B, 2H*2W, C] (bf16): the adjoint of the nearest-x2 replication."""
    if _is32(du):
        _req(du, f32, "du"); _req(dx, f32, "dx")
        check(lib().mfb_sumpool2x2_f32(_ptr(du), B, H, W, dx.shape[-1], _ptr(dx), _stream()))
        return
    _req(du, bf16, "du"); _req(dx, bf16, "dx")
    check(lib().mfb_sumpool2x2(_ptr(du), B, H, W, dx.shape[-1], _ptr(dx), _stream()))


def rowsum_per_image(dy, out, *, B, HW):
    """out [B, C] fp32 = per-image sums of dy [B, HW, C] over the pixels (gradient of the time-embedding row bias)."""
    _req(out, f32, "out")
    check(lib().mfb_rowsum_per_image(_ptr(dy), int(_is32(dy)), B, HW, dy.shape[-1], _ptr(out), _stream()))


def silu_bwd(x, dy=None, y=None, dx=None):
    """y = silu(x) and / or dx = dy * silu'(x); x, y, dx fp32, dy fp32 or bf16."""
    _req(x, f32, "x")
    check(lib().mfb_silu_bwd(_ptr(x), _ptr(dy), 1 if dy is None else int(_is32(dy)), _ptr(y), _ptr(dx), x.numel(), _stream()))


# fp32 parity-mode backward of attention / LayerNorm / GEGLU (the frozen UNet's data-gradient chain; not yet run on a GPU)
def attention_bwd_f32(q, k, v, d_out, dq, dk, dv, stats_ws, *, B, heads, head_dim, Tq, Tk):
    for name, t in (("q", q), ("k", k), ("v", v), ("d_out", d_out), ("dq", dq), ("dk", dk), ("dv", dv), ("stats_ws", stats_ws)):
        _req(t, f32, name)
    if stats_ws.numel() < 2 * B * heads * Tq:
        raise ValueError("attention_bwd_f32: stats workspace smaller than 2*B*heads*Tq floats")
    check(lib().mfb_attention_bwd_f32(_ptr(q), q.shape[-1], _ptr(k), k.shape[-1], _ptr(v), v.shape[-1], _ptr(d_out), d_out.shape[-1],
                                      _ptr(dq), dq.shape[-1], _ptr(dk), dk.shape[-1], _ptr(dv), dv.shape[-1], _ptr(stats_ws), B, heads,
                                      head_dim, Tq, Tk, _stream()))


def layernorm_bwd_f32(x, dy, gamma, dx, eps=1e-5):
    for name, t in (("x", x), ("dy", dy), ("gamma", gamma), ("dx", dx)):
        _req(t, f32, name)
    check(lib().mfb_layernorm_bwd_f32(_ptr(x), _ptr(dy), x.numel() // x.shape[-1], x.shape[-1], eps, _ptr(gamma), _ptr(dx), _stream()))


def geglu_f32(proj, out=None, d_out=None, d_proj=None):
    _req(proj, f32, "proj")
    Cc = proj.shape[-1] // 2
    check(lib().mfb_geglu_f32(_ptr(proj), proj.numel() // proj.shape[-1], Cc, _ptr(out), _ptr(d_out), _ptr(d_proj), _stream()))
