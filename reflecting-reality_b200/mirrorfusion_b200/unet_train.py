"""The FROZEN UNet of the BrushNet fine-tune step (BASELINE config 4): forward with every activation kept, and the data-gradient
chain that carries d loss / d model_pred back to the 28 BrushNet tap sites.

Reference: `MirrorFusionModel.forward` (E/train_brushnet_mirror.py:836-888) feeds the BrushNet taps into
`UNet2DConditionModel.forward` (S/models/unets/unet_2d_condition.py:1039-1348, tap adds at :1218,1289 and in the blocks,
S/models/unets/unet_2d_blocks.py:1388-1398,1483-1493,2626-2635,2751-2761); `accelerator.backward(loss)` (:1459) then runs autograd
through the whole UNet although `unet.requires_grad_(False)` (:1144): no UNet parameter gets a gradient, but every op's INPUT
gradient is needed, because a tap is added to a hidden state (`sample = sample + tap`), so d tap = d hidden at that site.

`FrozenUNetTrainer` is that program on the kernels — per op only the data gradient:
    conv / linear        the same tcgen05 implicit-GEMM plan on d y with the flipped / transposed weight (packed ONCE: frozen)
    GroupNorm(+SiLU)     ops.groupnorm_bwd (two-source for the skip concat; residual- and skip-path gradients added in the pass)
    LayerNorm            ops.layernorm_bwd (+ the transformer's residual gradient)
    GEGLU                ops.geglu on the kept projection
    attention            ops.attention_bwd (tcgen05 flash backward; cross attention: dq only — the text context is not trained)
    Upsample2D           stride-1 data-gradient plan at high resolution + 2x2 sum-pool;  Downsample2D: the sub-pixel plan
    conv_out             ops.conv_out_bwd from the fp32 NCHW loss gradient
A hidden state with two consumers (next block + skip) gets its gradient SUMMED inside the main consumer's last backward kernel
(`dres2` of the GroupNorm backward / `res1` of the downsampler's data-gradient plan), never by a separate add.  A tap is added to
a block's output, so d tap = the total gradient of that output: the tap gradients handed to `BrushNetTrainer.backward` are those
buffers themselves.

Written against a kernel namespace `K` (default: `ops`, the C-ABI library); tests/torch_kernels.py is the CPU stand-in used to
check the DATAFLOW against float64 autograd of the oracle UNet.  The product path never uses it (ops raises without CUDA).
"""
from __future__ import annotations

import os

from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import ops as _ops
from .config import NetConfig, tap_channels, up_block_channels
from .engine import _resnet_prefixes

f32 = torch.float32


def _dgrad_linear(w: torch.Tensor) -> torch.Tensor:
    """[N, K] linear / 1x1-conv weight -> weight of its data gradient ([K, N]: dx = dy W)."""
    if w.dim() == 4:
        w = w[:, :, 0, 0]
    return w.t().contiguous()


class FrozenUNetTrainer:
    def __init__(self, cfg: NetConfig, sd: Dict[str, torch.Tensor], taps: List[torch.Tensor], *, B: int, H: int, W: int, device,
                 ctx_len: int = 77, precision: str = "bf16", K=None):
        """precision="fp32": the PARITY MODE — the same program on the CUDA-core fp32 kernels (csrc/fp32mode.cu, train.cu), for the
        1e-3 bar against float64 autograd of the oracle on the device.
        taps: the 28 tap tensors [B, h*w, C] in the reference's pop order (12 down, mid, 15 up for SD1.5) — the BrushNet
        trainer's buffers, read in place by the producing epilogues (`res2`)."""
        K = _ops if K is None else K
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.K, self.cfg, self.B, self.H, self.W, self.dev, self.ctx_len = K, cfg, B, H, W, torch.device(device), ctx_len
        self.act = f32 if precision == "fp32" else torch.bfloat16
        self.sd = {k: v.detach().to(device=device, dtype=f32) for k, v in sd.items()}
        self.fwd: List[Callable[[], None]] = []
        self.bwd: List[Callable[[], None]] = []
        self.keep: List[object] = []
        self._scratch: Dict[Tuple, torch.Tensor] = {}
        self.flops_fwd = self.flops_bwd = 0.0
        with _ops.precision(precision):      # the weight-packing helpers read the storage dtype
            self._build(taps)

    def _build(self, taps):
        K, cfg, B, H, W, dev = self.K, self.cfg, self.B, self.H, self.W, self.dev
        boc, n = cfg.block_out_channels, len(cfg.block_out_channels)
        dch, _mch, uch = tap_channels(cfg)
        if len(taps) != len(dch) + 1 + len(uch):
            raise ValueError(f"expected {len(dch) + 1 + len(uch)} tap tensors, got {len(taps)}")
        self.gn_ws = torch.zeros(K.gn_ws_floats(B, cfg.norm_num_groups), device=dev, dtype=f32)
        self.sample_in = torch.zeros(B, cfg.in_channels, H, W, device=dev, dtype=f32)
        self.ehs_in = torch.zeros(B, self.ctx_len, cfg.cross_attention_dim, device=dev, dtype=f32)
        self.ehs_act = torch.zeros(B * self.ctx_len, cfg.cross_attention_dim, device=dev, dtype=self.act)
        self.out = torch.zeros(B, cfg.out_channels, H, W, device=dev, dtype=f32)        # model_pred, NCHW fp32
        self.d_pred = torch.zeros_like(self.out)                                         # d loss / d model_pred
        self.t_dev = torch.zeros(B, device=dev, dtype=f32)
        self.fwd.append(lambda: K.f32_to_bf16(self.ehs_in, self.ehs_act))
        self._time_path(_resnet_prefixes(cfg))

        tap_it = iter(taps)
        wci = self.sd["conv_in.weight"].permute(2, 3, 1, 0).contiguous()
        bci = self.sd["conv_in.bias"].contiguous()
        hw = (H, W)
        pre, x = self.buf(B, H * W, boc[0]), self.buf(B, H * W, boc[0])     # the first skip keeps the PRE-tap conv_in output (:1215-1218)
        tap0 = next(tap_it)
        self.keep += [wci, bci]
        self.fwd.append(lambda x0=x: K.conv_in(self.sample_in, None, wci, bci, pre, tap0, x0))      # bind now: `x` is reassigned below
        blocks: List[_Block] = []
        skips: List[Tuple[torch.Tensor, Tuple[int, int], int]] = [(pre, hw, -1)]         # (tensor, hw, index of the producing block)
        for i in range(n):
            for j in range(cfg.layers_per_block):
                tap = next(tap_it)
                if cfg.down_has_attn[i]:
                    r = _Resnet(self, f"down_blocks.{i}.resnets.{j}", x, None, hw, boc[i], None)
                    t = _Transformer(self, f"down_blocks.{i}.attentions.{j}", r.out, hw, tap)
                    blocks += [r, t]
                    x = t.out
                else:
                    r = _Resnet(self, f"down_blocks.{i}.resnets.{j}", x, None, hw, boc[i], tap)
                    blocks.append(r)
                    x = r.out
                skips.append((x, hw, len(blocks) - 1))
            if i != n - 1:
                d = _Downsample(self, f"down_blocks.{i}.downsamplers.0", x, hw, next(tap_it))
                blocks.append(d)
                x, hw = d.out, (hw[0] // 2, hw[1] // 2)
                skips.append((x, hw, len(blocks) - 1))
        r = _Resnet(self, "mid_block.resnets.0", x, None, hw, boc[-1], None)
        t = _Transformer(self, "mid_block.attentions.0", r.out, hw, None)
        r2 = _Resnet(self, "mid_block.resnets.1", t.out, None, hw, boc[-1], next(tap_it))
        blocks += [r, t, r2]
        x = r2.out
        skip_consumer: Dict[int, _Resnet] = {}      # producing block index -> the up resnet that takes its output as skip
        for i, layers in enumerate(up_block_channels(cfg)):
            for j, (_cin, _hid, _skip, cout) in enumerate(layers):
                s, shw, src = skips.pop()
                assert shw == hw
                tap = next(tap_it)
                if cfg.up_has_attn[i]:
                    r = _Resnet(self, f"up_blocks.{i}.resnets.{j}", x, s, hw, cout, None)
                    t = _Transformer(self, f"up_blocks.{i}.attentions.{j}", r.out, hw, tap)
                    blocks += [r, t]
                    x = t.out
                else:
                    r = _Resnet(self, f"up_blocks.{i}.resnets.{j}", x, s, hw, cout, tap)
                    blocks.append(r)
                    x = r.out
                skip_consumer[src] = r
            if i != n - 1:
                u = _Upsample(self, f"up_blocks.{i}.upsamplers.0", x, hw, next(tap_it))
                blocks.append(u)
                x, hw = u.out, (hw[0] * 2, hw[1] * 2)
        self.blocks = blocks
        self.head = _Head(self, x, hw)
        # ---- backward construction, in execution (reverse) order.  `d` = TOTAL gradient of blocks[idx].out.  Block idx's input is
        # blocks[idx-1].out; if that is also a skip, the up resnet that consumed it has already been built (it comes later in
        # forward order) and its skip-input gradient is folded into block idx's input gradient.
        d = self.head.build_backward()
        for idx in range(len(blocks) - 1, -1, -1):
            blocks[idx].d_out_total = d
            # (block 0's input is conv_in's POST-tap output; the skip taken at the conv_in site is the PRE-tap tensor, :1215-1218,
            # whose gradient reaches no trainable parameter)
            sc = skip_consumer.get(idx - 1) if idx > 0 else None
            d = blocks[idx].build_backward(d, None if sc is None else sc.dxb)
        self.d_taps: List[torch.Tensor] = [d] + [b.d_out_total for b in blocks if b.tap is not None]     # tap 0: the conv_in site
        assert len(self.d_taps) == len(taps)
        self.n_down = len(dch)

    # ------------------------------------------------------------------------------------------------ memory / weights
    def buf(self, *shape, dtype=None) -> torch.Tensor:
        t = torch.zeros(*shape, device=self.dev, dtype=dtype or self.act)
        self.keep.append(t)
        return t

    def scratch(self, role: str, *shape) -> torch.Tensor:
        """Backward temporaries that are consumed immediately: shared by all blocks of equal shape."""
        key = (role,) + tuple(shape)
        if key not in self._scratch:
            self._scratch[key] = torch.zeros(*shape, device=self.dev, dtype=self.act)
        return self._scratch[key]

    def scratch32(self, role: str, *shape) -> torch.Tensor:
        key = (role, "f32") + tuple(shape)
        if key not in self._scratch:
            self._scratch[key] = torch.zeros(*shape, device=self.dev, dtype=f32)
        return self._scratch[key]

    def w(self, name: str) -> torch.Tensor:
        return self.sd[name].contiguous()

    def gn_keep(self, HW: int):
        """A GroupNorm workspace of its own for one norm, so that its forward statistics ({sum, sum of squares} per image and
        group at the head of the workspace) survive until its backward: -> (workspace, what to pass as `stats` to groupnorm_bwd).
        Maps up to 8x8 go through the single-launch kernel, which keeps the statistics in registers: their backward recomputes."""
        ws = torch.zeros(self.K.gn_ws_floats(self.B, self.cfg.norm_num_groups), device=self.dev, dtype=f32)
        self.keep.append(ws)
        return ws, (ws if HW > 64 and os.environ.get("MFB_TRAIN_KEEP_GN_STATS", "1") == "1" else None)      # =0: recompute (A/B)

    def wa(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(self.act).contiguous()

    def plan(self, x, w, out, *, B, H, W, Cin, Cout, ksize=1, bwd=False, **kw):
        p = self.K.ConvPlan(x, w, out, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=ksize, **kw)
        self.keep.append(p)
        if bwd:
            self.flops_bwd += getattr(p, "flops", 0.0)
        else:
            self.flops_fwd += getattr(p, "flops", 0.0)
        return p

    def linear(self, x, w, out, bwd=False, **kw):
        M, Kd = x.shape
        return self.plan(x, w, out, B=1, H=1, W=M, Cin=Kd, Cout=w.shape[0], ksize=1, bwd=bwd, **kw)

    # ------------------------------------------------------------------------------------------------ timestep path (frozen)
    def _time_path(self, prefixes):
        K, cfg, B = self.K, self.cfg, self.B
        c0, temb = cfg.block_out_channels[0], cfg.time_embed_dim
        sin, e1, emb = self.buf(B, c0, dtype=f32), self.buf(B, temb, dtype=f32), self.buf(B, temb, dtype=f32)
        w1, b1 = self.wa(self.sd["time_embedding.linear_1.weight"]), self.w("time_embedding.linear_1.bias")
        w2, b2 = self.wa(self.sd["time_embedding.linear_2.weight"]), self.w("time_embedding.linear_2.bias")
        wcat = self.wa(torch.cat([self.sd[p + ".time_emb_proj.weight"] for p in prefixes], 0))
        bcat = torch.cat([self.sd[p + ".time_emb_proj.bias"] for p in prefixes], 0).contiguous()
        self.rowbias = self.buf(B, wcat.shape[0], dtype=f32)
        self.rowbias_off, off = {}, 0
        for p in prefixes:
            self.rowbias_off[p] = off
            off += self.sd[p + ".time_emb_proj.weight"].shape[0]
        self.keep += [w1, b1, w2, b2, wcat, bcat]
        self.fwd += [lambda: K.timestep_sinusoid(self.t_dev, sin), lambda: K.linear_small(sin, w1, b1, e1, act_out=True),
                     lambda: K.linear_small(e1, w2, b2, emb), lambda: K.linear_small(emb, wcat, bcat, self.rowbias, act_in=True)]

    # ------------------------------------------------------------------------------------------------ run
    def forward(self, sample: torch.Tensor, timesteps: torch.Tensor, encoder_hidden_states: torch.Tensor) -> torch.Tensor:
        """sample [B,4,H,W] (noisy latents), timesteps [B], encoder_hidden_states [B,77,ctx]; the taps are read from the tensors
        given at construction.  Returns model_pred [B,4,H,W] fp32 (a buffer owned by the trainer)."""
        self.sample_in.copy_(sample)
        self.t_dev.copy_(timesteps.to(f32))
        self.ehs_in.copy_(encoder_hidden_states.to(f32))
        for f in self.fwd:
            f()
        return self.out

    def backward(self, d_pred: Optional[torch.Tensor] = None):
        """d_pred: d loss / d model_pred [B,4,H,W] fp32 (None: already in self.d_pred, e.g. written by TrainLoss).  Returns the tap
        gradients (d_down[12], d_mid, d_up[15]) as buffers owned by the trainer, [B, h*w, C] each."""
        if d_pred is not None:
            self.d_pred.copy_(d_pred)
        for f in self.bwd:
            f()
        nd = self.n_down
        return self.d_taps[:nd], self.d_taps[nd], self.d_taps[nd + 1:]


class _Block:
    tap: Optional[torch.Tensor] = None
    d_out_total: Optional[torch.Tensor] = None       # the buffer holding the TOTAL gradient of this block's output (= d tap)
    out: torch.Tensor


class _Resnet(_Block):
    """ResnetBlock2D (S/models/resnet.py:329-405), frozen.  Forward keeps xa / xb (inputs) and c1 (pre-norm2)."""

    def __init__(self, T: FrozenUNetTrainer, p: str, xa, xb, hw, cout: int, tap):
        K, B = T.K, T.B
        self.T, self.p, self.xa, self.xb, self.hw, self.cout, self.tap = T, p, xa, xb, hw, cout, tap
        h, w = hw
        HW = h * w
        self.ca, self.cb = xa.shape[-1], (0 if xb is None else xb.shape[-1])
        cin = self.ca + self.cb
        n1, n2 = T.scratch("n1", B, HW, cin), T.scratch("n2", B, HW, cout)      # normalised inputs are not needed again (frozen: no wgrad)
        self.c1, self.out = T.buf(B, HW, cout), T.buf(B, HW, cout)
        self.gn = tuple(T.w(p + s) for s in (".norm1.weight", ".norm1.bias", ".norm2.weight", ".norm2.bias"))
        g1, b1, g2, b2 = self.gn
        off = T.rowbias_off[p]
        plan1 = T.plan(n1, _ops.pack_conv_weight(T.sd[p + ".conv1.weight"]), self.c1, B=B, H=h, W=w, Cin=cin, Cout=cout, ksize=3,
                       bias=T.w(p + ".conv1.bias"), rowbias=T.rowbias[:, off:], rowbias_ld=T.rowbias.shape[1])
        bias = T.sd[p + ".conv2.bias"].clone()
        extras_w, extras_x, res1 = [], [], None
        self.has_sc = p + ".conv_shortcut.weight" in T.sd
        if self.has_sc:      # the 1x1 shortcut over the (concatenated) input = extra K-segments of conv2's GEMM
            ws = T.sd[p + ".conv_shortcut.weight"][:, :, 0, 0]
            extras_w += [ws[:, :self.ca]] + ([ws[:, self.ca:]] if self.cb else [])
            extras_x += [xa] + ([xb] if self.cb else [])
            bias = bias + T.sd[p + ".conv_shortcut.bias"]
        else:
            assert self.cb == 0 and self.ca == cout
            res1 = xa
        plan2 = T.plan(n2, _ops.pack_conv_weight(T.sd[p + ".conv2.weight"], extras=extras_w), self.out, B=B, H=h, W=w, Cin=cout, Cout=cout,
                       ksize=3, extras=extras_x, bias=bias.contiguous(), res1=res1, res2=tap)
        gn = dict(B=B, HW=HW, groups=T.cfg.norm_num_groups, eps=T.cfg.norm_eps, silu=True)
        (ws1, self.st1), (ws2, self.st2) = T.gn_keep(HW), T.gn_keep(HW)
        T.fwd += [lambda: K.groupnorm(xa, xb, g1, b1, n1, ws1, **gn), plan1.run,
                  lambda: K.groupnorm(self.c1, None, g2, b2, n2, ws2, **gn), plan2.run]

    def build_backward(self, d_out, extra_in):
        """d_out: total gradient of self.out.  extra_in: gradient of the MAIN input arriving over the skip path (or None), added
        in the last kernel.  Returns the main input's gradient buffer (self.dxa); the skip input's is self.dxb."""
        T, K, B = self.T, self.T.K, self.T.B
        h, w = self.hw
        HW, cin, cout, p = h * w, self.ca + self.cb, self.cout, self.p
        g1, b1, g2, b2 = self.gn
        dn2, dc1, dn1 = T.scratch("dn2", B, HW, cout), T.scratch("dc1", B, HW, cout), T.scratch("dn1", B, HW, cin)
        self.dxa = T.buf(B, HW, self.ca)
        self.dxb = T.buf(B, HW, self.cb) if self.cb else None
        plan_d2 = T.plan(d_out, _ops.pack_conv_dgrad_weight(T.sd[p + ".conv2.weight"]), dn2, B=B, H=h, W=w, Cin=cout, Cout=cout, ksize=3, bwd=True)
        plan_d1 = T.plan(dc1, _ops.pack_conv_dgrad_weight(T.sd[p + ".conv1.weight"]), dn1, B=B, H=h, W=w, Cin=cout, Cout=cin, ksize=3, bwd=True)
        gnb = dict(B=B, HW=HW, groups=T.cfg.norm_num_groups, eps=T.cfg.norm_eps, silu=True)
        T.bwd += [plan_d2.run, lambda: K.groupnorm_bwd(self.c1, None, dn2, g2, b2, dc1, None, None, stats=self.st2, **gnb), plan_d1.run]
        if self.has_sc:      # shortcut path: its data gradient over the whole (concatenated) input, one 1x1 plan
            dsc = T.scratch("dsc", B, HW, cin)
            plan_dsc = T.plan(d_out, T.wa(_dgrad_linear(T.sd[p + ".conv_shortcut.weight"])), dsc, B=B, H=h, W=w, Cin=cout, Cout=cin, ksize=1, bwd=True)
            T.bwd.append(plan_dsc.run)
            dres = dsc
        else:
            dres = d_out     # identity shortcut (resnet.py:403)
        T.bwd.append(lambda: K.groupnorm_bwd(self.xa, self.xb, dn1, g1, b1, self.dxa, self.dxb, None, dres=dres, dres2=extra_in,
                                             stats=self.st1, **gnb))
        return self.dxa


class _Downsample(_Block):
    """Downsample2D: conv3x3 stride 2 (S/models/downsampling.py:134-154) + tap."""

    def __init__(self, T, p, x, hw, tap):
        B, (h, w), c = T.B, hw, x.shape[-1]
        self.T, self.p, self.x, self.hw, self.tap = T, p, x, hw, tap
        self.out = T.buf(B, (h // 2) * (w // 2), c)
        plan = T.plan(x, _ops.pack_conv_weight(T.sd[p + ".conv.weight"]), self.out, B=B, H=h, W=w, Cin=c, Cout=c, ksize=3, stride=2,
                      bias=T.w(p + ".conv.bias"), res2=tap)
        T.fwd.append(plan.run)

    def build_backward(self, d_out, extra_in):
        T, B, (h, w), c = self.T, self.T.B, self.hw, self.x.shape[-1]
        self.dx = T.buf(B, h * w, c)
        # the sub-pixel (`up2x`) plan over d y with parity-selected taps (ops.pack_conv_s2_dgrad_weight); the skip-path gradient of the
        # input rides in as res1
        plan = T.plan(d_out, _ops.pack_conv_s2_dgrad_weight(T.sd[self.p + ".conv.weight"]), self.dx, B=B, H=h // 2, W=w // 2, Cin=c, Cout=c,
                      ksize=3, up2x=True, res1=extra_in, bwd=True)
        T.bwd.append(plan.run)
        return self.dx


class _Upsample(_Block):
    """Upsample2D: nearest x2 + conv3x3 (S/models/upsampling.py:145-186) as the sub-pixel plan, + tap."""

    def __init__(self, T, p, x, hw, tap):
        B, (h, w), c = T.B, hw, x.shape[-1]
        self.T, self.p, self.x, self.hw, self.tap = T, p, x, hw, tap
        self.out = T.buf(B, 4 * h * w, c)
        plan = T.plan(x, _ops.pack_upconv_weight(T.sd[p + ".conv.weight"]), self.out, B=B, H=h, W=w, Cin=c, Cout=c, ksize=3, up2x=True,
                      bias=T.w(p + ".conv.bias"), res2=tap)
        T.fwd.append(plan.run)

    def build_backward(self, d_out, extra_in):
        assert extra_in is None
        T, K, B, (h, w), c = self.T, self.T.K, self.T.B, self.hw, self.x.shape[-1]
        du = T.scratch("du", B, 4 * h * w, c)
        self.dx = T.buf(B, h * w, c)
        plan = T.plan(d_out, _ops.pack_conv_dgrad_weight(T.sd[self.p + ".conv.weight"]), du, B=B, H=2 * h, W=2 * w, Cin=c, Cout=c, ksize=3, bwd=True)
        T.bwd += [plan.run, lambda: K.sumpool2x2(du, self.dx, B=B, H=h, W=w)]       # adjoint of the nearest x2 replication
        return self.dx


class _Transformer(_Block):
    """Transformer2DModel + BasicTransformerBlock (S/models/transformers/transformer_2d.py:334-430, S/models/attention.py:291-412),
    frozen.  Forward keeps: x (input), h0 (after proj_in), qkv, att, lse, h1, q2, k2, v2, att2, lse2, h2, the GEGLU projection."""

    def __init__(self, T: FrozenUNetTrainer, p: str, x, hw, tap):
        K, B, cfg = T.K, T.B, T.cfg
        self.T, self.p, self.x, self.hw, self.tap = T, p, x, hw, tap
        h, w = hw
        Tn = h * w
        M, C = B * Tn, x.shape[-1]
        heads = cfg.heads
        d = C // heads
        Lc = T.ctx_len
        t = p + ".transformer_blocks.0"
        self.t = t
        sd = T.sd
        g = T.scratch("tg", B, Tn, C)
        nrm = T.scratch("tn", M, C)
        self.h0, self.qkv, self.att, self.h1 = T.buf(M, C), T.buf(M, 3 * C), T.buf(M, C), T.buf(M, C)
        self.q2, self.k2, self.v2, self.att2, self.h2 = T.buf(M, C), T.buf(B * Lc, C), T.buf(B * Lc, C), T.buf(M, C), T.buf(M, C)
        self.proj, self.out = T.buf(M, 8 * C), T.buf(B, Tn, C)
        gg, h3 = T.scratch("tgg", M, 4 * C), T.scratch("th3", M, C)
        self.lse, self.lse2 = T.buf(B * heads * Tn, dtype=f32), T.buf(B * heads * Tn, dtype=f32)
        self.gn = (T.w(p + ".norm.weight"), T.w(p + ".norm.bias"))
        self.ln = [(T.w(f"{t}.norm{i}.weight"), T.w(f"{t}.norm{i}.bias")) for i in (1, 2, 3)]
        gnw, gnb_ = self.gn
        wa = T.wa
        p_in = T.linear(g.view(M, C), _ops.pack_conv_weight(sd[p + ".proj_in.weight"]), self.h0, bias=T.w(p + ".proj_in.bias"))
        wqkv = wa(torch.cat([sd[t + ".attn1.to_q.weight"], sd[t + ".attn1.to_k.weight"], sd[t + ".attn1.to_v.weight"]], 0))
        p_qkv = T.linear(nrm, wqkv, self.qkv)
        p_o1 = T.linear(self.att, wa(sd[t + ".attn1.to_out.0.weight"]), self.h1, bias=T.w(t + ".attn1.to_out.0.bias"), res1=self.h0)
        p_k2 = T.linear(T.ehs_act, wa(sd[t + ".attn2.to_k.weight"]), self.k2)
        p_v2 = T.linear(T.ehs_act, wa(sd[t + ".attn2.to_v.weight"]), self.v2)
        p_q2 = T.linear(nrm, wa(sd[t + ".attn2.to_q.weight"]), self.q2)
        p_o2 = T.linear(self.att2, wa(sd[t + ".attn2.to_out.0.weight"]), self.h2, bias=T.w(t + ".attn2.to_out.0.bias"), res1=self.h1)
        # GEGLU un-fused: the projection [h | gate] is kept for the backward (activations.py:100-103)
        p_ff1 = T.linear(nrm, wa(sd[t + ".ff.net.0.proj.weight"]), self.proj, bias=T.w(t + ".ff.net.0.proj.bias"))
        p_ff2 = T.linear(gg, wa(sd[t + ".ff.net.2.weight"]), h3, bias=T.w(t + ".ff.net.2.bias"), res1=self.h2)
        p_out = T.plan(h3, _ops.pack_conv_weight(sd[p + ".proj_out.weight"]), self.out, B=B, H=h, W=w, Cin=C, Cout=C, ksize=1,
                       bias=T.w(p + ".proj_out.bias"), res1=x, res2=tap)
        (l1g, l1b), (l2g, l2b), (l3g, l3b) = self.ln
        qkv = self.qkv
        kview, vview = qkv.view(-1)[C:], qkv.view(-1)[2 * C:]
        self.geo = dict(B=B, heads=heads, head_dim=d, Tq=Tn)
        ws_gn, self.st = T.gn_keep(Tn)
        T.fwd += [
            lambda: K.groupnorm(x, None, gnw, gnb_, g, ws_gn, B=B, HW=Tn, groups=cfg.norm_num_groups, eps=1e-6, silu=False),
            p_in.run,
            lambda: K.layernorm(self.h0, l1g, l1b, nrm, 1e-5), p_qkv.run,
            lambda: K.attention_lse(qkv, kview, vview, self.att, self.lse, Tk=Tn, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, **self.geo),
            p_o1.run,
            p_k2.run, p_v2.run,
            lambda: K.layernorm(self.h1, l2g, l2b, nrm, 1e-5), p_q2.run,
            lambda: K.attention_lse(self.q2, self.k2, self.v2, self.att2, self.lse2, Tk=Lc, ldq=C, ldk=C, ldv=C, ldo=C, **self.geo),
            p_o2.run,
            lambda: K.layernorm(self.h2, l3g, l3b, nrm, 1e-5), p_ff1.run,
            lambda: K.geglu(self.proj, out=gg),
            p_ff2.run, p_out.run]
        T.flops_fwd += 4.0 * B * Tn * (Tn + Lc) * C

    def build_backward(self, d_out, extra_in):
        assert extra_in is None        # a transformer's input (a resnet output) is never a skip
        T, K, B, cfg = self.T, self.T.K, self.T.B, self.T.cfg
        h, w = self.hw
        Tn = h * w
        M, C = B * Tn, self.x.shape[-1]
        Lc, heads, t, p, sd = T.ctx_len, cfg.heads, self.t, self.p, T.sd
        wa = T.wa
        S = T.scratch
        dh3, dgg, dproj, dn = S("dh3", M, C), S("dgg", M, 4 * C), S("dproj", M, 8 * C), S("dnrm", M, C)
        dh2, datt, dq2, dh1, dqkv, dh0, dg = S("dh2", M, C), S("datt", M, C), S("dq2", M, C), S("dh1", M, C), S("dqkv", M, 3 * C), S("dh0", M, C), S("dg", B, Tn, C)
        dvec = T.scratch32("dvec", B * heads * Tn)
        self.dx = T.buf(B, Tn, C)
        d_out2 = d_out.view(M, C)
        lin = lambda x_, w_, o_, **kw: T.linear(x_, wa(_dgrad_linear(w_)), o_, bwd=True, **kw)
        b_out = lin(d_out2, sd[p + ".proj_out.weight"], dh3)                                   # d h3;  d x (residual) = d_out
        b_ff2 = lin(dh3, sd[t + ".ff.net.2.weight"], dgg)                                      # d gg;  d h2 (residual) = d h3
        b_ff1 = lin(dproj, sd[t + ".ff.net.0.proj.weight"], dn)                                # d norm3(h2)
        b_o2 = lin(dh2, sd[t + ".attn2.to_out.0.weight"], datt)                                # d att2;  d h1 (residual) = d h2
        b_q2 = lin(dq2, sd[t + ".attn2.to_q.weight"], dn)                                      # d norm2(h1)
        b_o1 = lin(dh1, sd[t + ".attn1.to_out.0.weight"], datt)                                # d att;  d h0 (residual) = d h1
        wqkv = torch.cat([sd[t + ".attn1.to_q.weight"], sd[t + ".attn1.to_k.weight"], sd[t + ".attn1.to_v.weight"]], 0)
        b_qkv = lin(dqkv, wqkv, dn)                                                            # d norm1(h0)
        b_in = lin(dh0, sd[p + ".proj_in.weight"], dg.view(M, C))                              # d groupnorm(x)
        (l1g, _), (l2g, _), (l3g, _) = self.ln
        gnw, gnb_ = self.gn
        qkv = self.qkv
        kview, vview = qkv.view(-1)[C:], qkv.view(-1)[2 * C:]
        dkview, dvview = dqkv.view(-1)[C:], dqkv.view(-1)[2 * C:]
        geo = self.geo
        T.bwd += [
            b_out.run, b_ff2.run,
            lambda: K.geglu(self.proj, d_out=dgg, d_proj=dproj),
            b_ff1.run,
            lambda: K.layernorm_bwd(self.h2, dn, l3g, dh2, 1e-5, dres=dh3),                    # total d h2
            b_o2.run,
            lambda: K.attention_bwd(self.q2, self.k2, self.v2, self.att2, datt, self.lse2, dvec, dq2, None, None, Tk=Lc, ldq=C, ldk=C,
                                    ldv=C, ldo=C, lddo=C, lddq=C, **geo),                       # frozen text context: dq only
            b_q2.run,
            lambda: K.layernorm_bwd(self.h1, dn, l2g, dh1, 1e-5, dres=dh2),                    # total d h1
            b_o1.run,
            lambda: K.attention_bwd(qkv, kview, vview, self.att, datt, self.lse, dvec, dqkv, dkview, dvview, Tk=Tn, ldq=3 * C, ldk=3 * C,
                                    ldv=3 * C, ldo=C, lddo=C, lddq=3 * C, lddk=3 * C, lddv=3 * C, **geo),
            b_qkv.run,
            lambda: K.layernorm_bwd(self.h0, dn, l1g, dh0, 1e-5, dres=dh1),                    # total d h0
            b_in.run,
            lambda: K.groupnorm_bwd(self.x, None, dg, gnw, gnb_, self.dx, None, None, B=B, HW=Tn, groups=cfg.norm_num_groups, eps=1e-6,
                                    silu=False, dres=d_out, stats=self.st)]                     # + the transformer's residual
        T.flops_bwd += 2.5 * 4.0 * B * Tn * Tn * C + 1.5 * 4.0 * B * Tn * Lc * C
        return self.dx


class _Head:
    """conv_norm_out -> SiLU -> conv_out (S/models/unets/unet_2d_condition.py:1336-1339)."""

    def __init__(self, T: FrozenUNetTrainer, x, hw):
        K, B, cfg = T.K, T.B, T.cfg
        self.T, self.x, self.hw = T, x, hw
        h, w = hw
        c0 = cfg.block_out_channels[0]
        nout = T.scratch("n1", B, h * w, c0)
        self.g, self.b = T.w("conv_norm_out.weight"), T.w("conv_norm_out.bias")
        self.wco = T.sd["conv_out.weight"].permute(0, 2, 3, 1).contiguous()       # [Cout, 3, 3, Cin] fp32
        bco = T.w("conv_out.bias")
        g, b, wco = self.g, self.b, self.wco
        T.keep += [wco, bco]
        ws_gn, self.st = T.gn_keep(h * w)
        T.fwd += [lambda: K.groupnorm(x, None, g, b, nout, ws_gn, B=B, HW=h * w, groups=cfg.norm_num_groups, eps=cfg.norm_eps, silu=True),
                  lambda: K.conv_out(nout, wco, bco, T.out, B=B, H=h, W=w)]

    def build_backward(self):
        T, K, B, cfg = self.T, self.T.K, self.T.B, self.T.cfg
        h, w = self.hw
        c0 = cfg.block_out_channels[0]
        dn = T.scratch("dn1", B, h * w, c0)
        self.dx = T.buf(B, h * w, c0)
        g, b, wco = self.g, self.b, self.wco
        T.bwd += [lambda: K.conv_out_bwd(T.d_pred, wco, dn, B=B, H=h, W=w),
                  lambda: K.groupnorm_bwd(self.x, None, dn, g, b, self.dx, None, None, B=B, HW=h * w, groups=cfg.norm_num_groups,
                                          eps=cfg.norm_eps, silu=True, stats=self.st)]
        return self.dx
