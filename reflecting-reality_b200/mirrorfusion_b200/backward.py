"""Forward-for-training and backward of one ResnetBlock2D (S/models/resnet.py:329-405) as a launch program over the kernels —
the unit the BrushNet branch is made of (22 per net; BASELINE config 4).  First piece of the
backward PROGRAM of the nets; the net-level program (saved activations of all blocks, the samplers, the frozen UNet's
dgrad-only chain) is not built yet.

Block boundary: (x [B, HW, Cin], rowbias [B, Cout] = time_emb_proj(silu(emb))) -> out [B, HW, Cout]; backward takes d out and
returns (d x, d rowbias) and ACCUMULATES the parameter gradients into the flat gradient buffer.  The timestep MLP and
`time_emb_proj` themselves live in the hoisted timestep path (engine.build_time_path); their (GEMV-sized) backward consumes
d rowbias and is part of the net-level program.

Parameters live in a `train.FlatParams` in the layout the kernels consume — conv weights PACKED [Cout, kh*kw*Cin] — so that
    forward        reads   flat.w(name) (bf16 working copy) / flat.p(name) (fp32 parity mode)
    weight grad    writes  flat.g(name) in the same packed order (ops.conv_wgrad)
    AdamW          updates master and working copy in one launch (train.B200AdamW)
with no per-step repacking except the flipped / transposed copies the data-gradient plans read (`refresh_dgrad_weights`).

`K` is the kernel namespace (default: `ops`, the C-ABI library).  tests/torch_kernels.py provides a CPU stand-in with the same
call signatures so that the program's dataflow is checked against the reference's autograd without a GPU; the product path
never uses it (ops raises without CUDA)."""
from __future__ import annotations

import os

from typing import Dict, Optional, Tuple

import torch

from . import ops as _ops
from .train import FlatParams


def resnet_param_shapes(prefix: str, Cin: int, Cout: int) -> Dict[str, Tuple[int, ...]]:
    """Flat-buffer entries of one block, conv weights in packed [Cout, k*k*Cin] layout."""
    s = {f"{prefix}.norm1.weight": (Cin,), f"{prefix}.norm1.bias": (Cin,),
         f"{prefix}.conv1.weight": (Cout, 9 * Cin), f"{prefix}.conv1.bias": (Cout,),
         f"{prefix}.norm2.weight": (Cout,), f"{prefix}.norm2.bias": (Cout,),
         f"{prefix}.conv2.weight": (Cout, 9 * Cout), f"{prefix}.conv2.bias": (Cout,)}
    if Cin != Cout:
        s[f"{prefix}.conv_shortcut.weight"] = (Cout, Cin)
        s[f"{prefix}.conv_shortcut.bias"] = (Cout,)
    return s


def pack_resnet_state_dict(prefix: str, sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference state_dict entries (OIHW) of the block -> fp32 tensors in the flat-buffer layout."""
    out = {}
    for k, v in sd.items():
        if not k.startswith(prefix + ".") or ".time_emb_proj." in k:
            continue
        v = v.float()
        out[k] = v.permute(0, 2, 3, 1).reshape(v.shape[0], -1).contiguous() if v.dim() == 4 else v.contiguous()
    return out


def unpack_conv_grad(g: torch.Tensor, ksize: int) -> torch.Tensor:
    """Packed [Cout, k*k*Cin] gradient / weight -> OIHW (for comparison with the reference's .grad)."""
    Cout = g.shape[0]
    return g.view(Cout, ksize, ksize, -1).permute(0, 3, 1, 2).contiguous()


def _dgrad_from_packed(wp: torch.Tensor, ksize: int, dtype) -> torch.Tensor:
    """Packed forward weight [Cout, k*k*Cin] -> packed data-gradient weight [Cin, k*k*Cout] (flip taps, swap channels)."""
    Cout = wp.shape[0]
    w = wp.view(Cout, ksize, ksize, -1).flip(1, 2).permute(3, 1, 2, 0)
    return w.reshape(w.shape[0], -1).to(dtype).contiguous()


class ResnetBlockTrainer:
    """`frozen=True`: data gradients only (d x, d rowbias) — a block of the frozen UNet, whose backward only has to carry the
    gradient to the BrushNet taps; no weight / affine gradients are computed or written."""

    def __init__(self, flat: FlatParams, prefix: str, *, B: int, H: int, W: int, Cin: int, Cout: int, groups: int = 32,
                 eps: float = 1e-5, precision: str = "bf16", K=None, frozen: bool = False):
        K = _ops if K is None else K
        self.K, self.flat, self.p, self.frozen = K, flat, prefix, frozen
        self.B, self.H, self.W, self.HW, self.Cin, self.Cout, self.groups, self.eps = B, H, W, H * W, Cin, Cout, groups, eps
        self.dt = torch.float32 if precision == "fp32" else torch.bfloat16
        self.shortcut = Cin != Cout
        dev = flat.param.device
        act = lambda c: torch.zeros(B, self.HW, c, device=dev, dtype=self.dt)
        # forward buffers (kept for the backward pass) and gradient buffers, allocated once: fixed addresses
        self.x, self.n1, self.c1, self.n2, self.out = act(Cin), act(Cin), act(Cout), act(Cout), act(Cout)
        self.sc = act(Cout) if self.shortcut else None
        self.rowbias = torch.zeros(B, Cout, device=dev, dtype=torch.float32)
        self.d_out, self.dn2, self.dc1, self.dn1, self.dx = act(Cout), act(Cout), act(Cout), act(Cin), act(Cin)
        self.dsc = act(Cin) if self.shortcut else None
        self.d_rowbias = torch.zeros(B, Cout, device=dev, dtype=torch.float32)
        # one GroupNorm workspace per norm: the forward leaves its per-(image, group) {sum, sum of squares} at the head of it, and
        # the backward reads them there instead of re-reading the tensor for a statistics pass (maps above 8x8: the single-launch
        # kernel of the small maps keeps its statistics in registers)
        self.gn_ws = torch.zeros(K.gn_ws_floats(B, groups), device=dev, dtype=torch.float32)
        self.gn_ws2 = torch.zeros(K.gn_ws_floats(B, groups), device=dev, dtype=torch.float32)
        self.keep_stats = self.HW > 64 and os.environ.get("MFB_TRAIN_KEEP_GN_STATS", "1") == "1"      # =0: recompute in the backward (A/B)
        self.gnb_ws = torch.zeros(2 * B * max(Cin, Cout), device=dev, dtype=torch.float32)
        self._wsrc = flat.p if self.dt == torch.float32 else flat.w        # fp32 masters (parity mode) or the bf16 working copy
        # data-gradient weights: the same implicit-GEMM kernel over the incoming gradient with flipped / transposed weights
        self.w1d = torch.zeros(Cin, 9 * Cout, device=dev, dtype=self.dt)
        self.w2d = torch.zeros(Cout, 9 * Cout, device=dev, dtype=self.dt)
        if self.shortcut:
            self.wscd = torch.zeros(Cin, Cout, device=dev, dtype=self.dt)
        self._make_plans()
        self.refresh_dgrad_weights()

    def _make_plans(self):
        K, flat, wsrc = self.K, self.flat, self._wsrc
        Cin, Cout = self.Cin, self.Cout
        n = lambda s: f"{self.p}.{s}"
        geo = dict(B=self.B, H=self.H, W=self.W)
        self.plan1 = K.ConvPlan(self.n1, wsrc(n("conv1.weight")), self.c1, Cin=Cin, Cout=Cout, ksize=3, bias=flat.p(n("conv1.bias")),
                                rowbias=self.rowbias, rowbias_ld=Cout, **geo)
        if self.shortcut:
            self.plan_sc = K.ConvPlan(self.x, wsrc(n("conv_shortcut.weight")), self.sc, Cin=Cin, Cout=Cout, ksize=1,
                                      bias=flat.p(n("conv_shortcut.bias")), **geo)
        self.plan2 = K.ConvPlan(self.n2, wsrc(n("conv2.weight")), self.out, Cin=Cout, Cout=Cout, ksize=3, bias=flat.p(n("conv2.bias")),
                                res1=self.sc if self.shortcut else self.x, **geo)
        self.plan_d2 = K.ConvPlan(self.d_out, self.w2d, self.dn2, Cin=Cout, Cout=Cout, ksize=3, **geo)
        self.plan_d1 = K.ConvPlan(self.dc1, self.w1d, self.dn1, Cin=Cout, Cout=Cin, ksize=3, **geo)
        if self.shortcut:
            self.plan_dsc = K.ConvPlan(self.d_out, self.wscd, self.dsc, Cin=Cout, Cout=Cin, ksize=1, **geo)

    def rebind(self, x=None, d_out=None):
        """Use the producer's output buffer as this block's input (and the consumer's input-gradient buffer as its output
        gradient) instead of private copies: forward() / backward() then skip the copy when handed exactly that tensor."""
        if x is not None:
            self.x = x.view_as(self.x)
        if d_out is not None:
            self.d_out = d_out.view_as(self.d_out)
        self._make_plans()

    def refresh_dgrad_weights(self):
        """Re-derive the data-gradient weights from the current parameters (after every optimizer step): a permuting copy."""
        n = lambda s: f"{self.p}.{s}"
        R = self.K.dgrad_repack
        R(self._wsrc(n("conv1.weight")), self.w1d, 3)
        R(self._wsrc(n("conv2.weight")), self.w2d, 3)
        if self.shortcut:
            R(self._wsrc(n("conv_shortcut.weight")), self.wscd, 1)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, rowbias: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B, HW, Cin] (activation dtype), rowbias [B, Cout] fp32 (None = 0).  Returns the block's output buffer."""
        K, f, n = self.K, self.flat, (lambda s: f"{self.p}.{s}")
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x.view_as(self.x))
        if rowbias is None:
            self.rowbias.zero_()
        else:
            self.rowbias.copy_(rowbias)
        gn = dict(B=self.B, HW=self.HW, groups=self.groups, eps=self.eps, silu=True)
        K.groupnorm(self.x, None, f.p(n("norm1.weight")), f.p(n("norm1.bias")), self.n1, self.gn_ws, **gn)     # resnet.py:337-338
        self.plan1.run()                                                                                      # :367 + :369-379
        K.groupnorm(self.c1, None, f.p(n("norm2.weight")), f.p(n("norm2.bias")), self.n2, self.gn_ws2, **gn)    # :381,393
        if self.shortcut:
            self.plan_sc.run()                                                                                # :398-401
        self.plan2.run()                                                                                      # :396 + :403
        return self.out

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, d_out: torch.Tensor):
        """d_out [B, HW, Cout].  Accumulates parameter gradients into flat.g(...); returns (d x, d rowbias) buffers."""
        K, f, n = self.K, self.flat, (lambda s: f"{self.p}.{s}")
        B, H, W, HW = self.B, self.H, self.W, self.HW
        if d_out.data_ptr() != self.d_out.data_ptr():
            self.d_out.copy_(d_out.view_as(self.d_out))
        gnb = dict(B=B, HW=HW, groups=self.groups, eps=self.eps, silu=True, accumulate=True)
        # out = conv2(n2) + bias2 + shortcut(x)
        tr = not self.frozen
        G = (lambda name: f.g(n(name))) if tr else (lambda name: None)
        if tr:
            K.conv_wgrad(self.n2, self.d_out, f.g(n("conv2.weight")), f.g(n("conv2.bias")), B=B, H=H, W=W, ksize=3, accumulate=True)
        self.plan_d2.run()                                                          # d n2
        # n2 = silu(groupnorm(c1))
        K.groupnorm_bwd(self.c1, None, self.dn2, f.p(n("norm2.weight")), f.p(n("norm2.bias")), self.dc1, None, self.gnb_ws,
                        dgamma=G("norm2.weight"), dbeta=G("norm2.bias"), stats=self.gn_ws2 if self.keep_stats else None, **gnb)
        # c1 = conv1(n1) + bias1 + rowbias[:, :, None, None]
        if tr:
            K.conv_wgrad(self.n1, self.dc1, f.g(n("conv1.weight")), f.g(n("conv1.bias")), B=B, H=H, W=W, ksize=3, accumulate=True)
        K.rowsum_per_image(self.dc1, self.d_rowbias, B=B, HW=HW)
        self.plan_d1.run()                                                          # d n1
        # shortcut path: identity -> d out itself; 1x1 conv -> its data gradient (+ its weight gradient)
        if self.shortcut:
            if tr:
                K.conv_wgrad(self.x, self.d_out, f.g(n("conv_shortcut.weight")), f.g(n("conv_shortcut.bias")), B=B, H=H, W=W, ksize=1,
                             accumulate=True)
            self.plan_dsc.run()
            dres = self.dsc
        else:
            dres = self.d_out
        # n1 = silu(groupnorm(x)); d x = that + the shortcut path's gradient, in the same pass
        K.groupnorm_bwd(self.x, None, self.dn1, f.p(n("norm1.weight")), f.p(n("norm1.bias")), self.dx, None, self.gnb_ws,
                        dgamma=G("norm1.weight"), dbeta=G("norm1.bias"), dres=dres, stats=self.gn_ws if self.keep_stats else None, **gnb)
        return self.dx, self.d_rowbias


class DownsampleTrainer:
    """Downsample2D (conv3x3, stride 2, padding 1; S/models/downsampling.py:134-154) forward and backward on the kernels.
    Forward: the stride-2 implicit-GEMM plan (four parity-view tensor maps).  Backward: weight / bias gradient by the stride-2
    weight-gradient kernel, data gradient by the sub-pixel `up2x` plan over d y with parity-selected taps
    (ops.pack_conv_s2_dgrad_weight) — the same tcgen05 kernel again.  Flat entries: `<prefix>.conv.weight` packed [C, 9*C],
    `<prefix>.conv.bias` [C]."""

    def __init__(self, flat: FlatParams, prefix: str, *, B: int, H: int, W: int, C: int, precision: str = "bf16", K=None,
                 frozen: bool = False):
        K = _ops if K is None else K
        if H % 2 or W % 2:
            raise ValueError("Downsample2D backward needs even H, W")
        self.K, self.flat, self.p, self.B, self.H, self.W, self.C, self.frozen = K, flat, prefix, B, H, W, C, frozen
        self.dt = torch.float32 if precision == "fp32" else torch.bfloat16
        dev = flat.param.device
        self._wsrc = flat.p if self.dt == torch.float32 else flat.w
        self.x = torch.zeros(B, H * W, C, device=dev, dtype=self.dt)
        self.out = torch.zeros(B, (H // 2) * (W // 2), C, device=dev, dtype=self.dt)
        self.d_out, self.dx = torch.zeros_like(self.out), torch.zeros_like(self.x)
        self.wd = torch.zeros(4, C, 4 * C, device=dev, dtype=self.dt)
        self._make_plans()
        self.refresh_dgrad_weights()

    def _make_plans(self):
        K, B, H, W, C = self.K, self.B, self.H, self.W, self.C
        self.plan = K.ConvPlan(self.x, self._wsrc(f"{self.p}.conv.weight"), self.out, B=B, H=H, W=W, Cin=C, Cout=C, ksize=3, stride=2,
                               bias=self.flat.p(f"{self.p}.conv.bias"))
        self.plan_d = K.ConvPlan(self.d_out, self.wd, self.dx, B=B, H=H // 2, W=W // 2, Cin=C, Cout=C, ksize=3, up2x=True)

    def rebind(self, x=None, d_out=None):
        if x is not None:
            self.x = x.view_as(self.x)
        if d_out is not None:
            self.d_out = d_out.view_as(self.d_out)
        self._make_plans()

    def refresh_dgrad_weights(self):
        w = unpack_conv_grad(self._wsrc(f"{self.p}.conv.weight").float(), 3)          # packed -> OIHW
        k_of = [[None, 1], [2, 0]]          # see ops.pack_conv_s2_dgrad_weight
        for py in range(2):
            for px in range(2):
                for ty in range(2):
                    for tx in range(2):
                        kh, kw = k_of[py][ty], k_of[px][tx]
                        blk = self.wd[py * 2 + px].view(self.C, 4, self.C)[:, ty * 2 + tx]
                        if kh is None or kw is None:
                            blk.zero_()
                        else:
                            blk.copy_(w[:, :, kh, kw].t().to(self.dt))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x.view_as(self.x))
        self.plan.run()
        return self.out

    def backward(self, d_out: torch.Tensor) -> torch.Tensor:
        f = self.flat
        if d_out.data_ptr() != self.d_out.data_ptr():
            self.d_out.copy_(d_out.view_as(self.d_out))
        if not self.frozen:     # frozen (UNet): data gradient only
            self.K.conv_wgrad(self.x, self.d_out, f.g(f"{self.p}.conv.weight"), f.g(f"{self.p}.conv.bias"), B=self.B, H=self.H, W=self.W,
                              ksize=3, stride=2, accumulate=True)
        self.plan_d.run()
        return self.dx


class ZeroConvTap:
    """One BrushNet zero-conv (1x1, S/models/brushnet.py:831-834,851): tap = conv1x1(h) (+ bias), forward and backward.
    Backward also folds the fan-out of h: d h = W^T d tap + (the gradient arriving from h's other consumer), the sum done by
    the data-gradient plan's residual epilogue.  Flat entries `<name>.weight` [C, C] (a packed 1x1) and `<name>.bias` [C]."""

    def __init__(self, flat: FlatParams, name: str, h: torch.Tensor, d_other: Optional[torch.Tensor], *, B, H, W, C, dt, K,
                 d_other2: Optional[torch.Tensor] = None):
        self.K, self.flat, self.name, self.h, self.geo = K, flat, name, h, dict(B=B, H=H, W=W)
        dev = flat.param.device
        self._wsrc = flat.p if dt == torch.float32 else flat.w
        self.tap = torch.zeros(B, H * W, C, device=dev, dtype=dt)
        self.d_tap, self.dh = torch.zeros_like(self.tap), torch.zeros_like(self.tap)
        self.wd = torch.zeros(C, C, device=dev, dtype=dt)
        self.C, self._others = C, (d_other, d_other2)
        self.plan = K.ConvPlan(h, self._wsrc(f"{name}.weight"), self.tap, Cin=C, Cout=C, ksize=1, bias=flat.p(f"{name}.bias"), **self.geo)
        self._make_plan_d()
        self.refresh_dgrad_weights()

    def _make_plan_d(self):
        # up to two more consumers of h (the next block; the up-block resnet that takes h as its skip): both residual inputs
        self.plan_d = self.K.ConvPlan(self.d_tap, self.wd, self.dh, Cin=self.C, Cout=self.C, ksize=1, res1=self._others[0],
                                      res2=self._others[1], **self.geo)

    def bind_d_tap(self, buf: torch.Tensor):
        """Read the tap gradient where the frozen UNet's backward leaves it (FrozenUNetTrainer.d_taps) instead of a private copy."""
        self.d_tap = buf.view_as(self.d_tap)
        self._make_plan_d()

    def refresh_dgrad_weights(self):
        self.K.dgrad_repack(self._wsrc(f"{self.name}.weight"), self.wd, 1)

    def forward(self):
        self.plan.run()
        return self.tap

    def backward(self, d_tap: torch.Tensor) -> torch.Tensor:
        f = self.flat
        if d_tap.data_ptr() != self.d_tap.data_ptr():
            self.d_tap.copy_(d_tap.view_as(self.d_tap))
        self.K.conv_wgrad(self.h, self.d_tap, f.g(f"{self.name}.weight"), f.g(f"{self.name}.bias"), ksize=1, accumulate=True, **self.geo)
        self.plan_d.run()
        return self.dh


def brushnet_down_mid_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """Flat-buffer entries (packed layouts) of BrushNet's down path, mid block and their 13 zero-convs
    (S/models/brushnet.py:296-372,831-851) — everything between conv_in_condition's output and the mid tap."""
    boc = cfg.block_out_channels
    s: Dict[str, Tuple[int, ...]] = {}
    cin = boc[0]
    k = 0
    s[f"brushnet_down_blocks.{k}.weight"], s[f"brushnet_down_blocks.{k}.bias"] = (cin, cin), (cin,)
    for i, cout in enumerate(boc):
        for j in range(cfg.layers_per_block):
            s.update(resnet_param_shapes(f"down_blocks.{i}.resnets.{j}", cin, cout))
            cin = cout
            k += 1
            s[f"brushnet_down_blocks.{k}.weight"], s[f"brushnet_down_blocks.{k}.bias"] = (cout, cout), (cout,)
        if i != len(boc) - 1:
            s[f"down_blocks.{i}.downsamplers.0.conv.weight"], s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (cout, 9 * cout), (cout,)
            k += 1
            s[f"brushnet_down_blocks.{k}.weight"], s[f"brushnet_down_blocks.{k}.bias"] = (cout, cout), (cout,)
    for j in range(2):
        s.update(resnet_param_shapes(f"mid_block.resnets.{j}", boc[-1], boc[-1]))
    s["brushnet_mid_block.weight"], s["brushnet_mid_block.bias"] = (boc[-1], boc[-1]), (boc[-1],)
    return s


def pack_brushnet_down_mid(cfg, sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference state_dict (OIHW) -> fp32 tensors in the layout of brushnet_down_mid_shapes."""
    out = {}
    for name, shp in brushnet_down_mid_shapes(cfg).items():
        v = sd[name].float()
        out[name] = (v.permute(0, 2, 3, 1).reshape(v.shape[0], -1) if v.dim() == 4 else v).reshape(shp).contiguous()
    return out


class BrushNetDownMidTrainer:
    """Forward-for-training and backward of the BrushNet branch from conv_in_condition's output to the mid tap: the down blocks
    (resnets + downsamplers), MidBlock2D, and the 12 + 1 zero-conv taps (oracle/mf_oracle.py `brushnet_forward`, first half;
    S/models/brushnet.py:810-851).  Net-level composition of ResnetBlockTrainer / DownsampleTrainer / ZeroConvTap: every hidden
    state feeds the next block AND its zero-conv, so in backward its gradient is the sum of the two — formed by the zero-conv's
    data-gradient plan (residual epilogue), never by a separate add.  conditioning_scale is 1 as in the reference's training
    script.  Not covered yet: conv_in_condition itself (boundary kernel), the up blocks (two-source resnets, Upsample2D backward),
    the timestep path (consumes the d rowbias this class returns)."""

    def __init__(self, flat: FlatParams, cfg, *, B: int, H: int, W: int, precision: str = "bf16", K=None):
        K = _ops if K is None else K
        self.K, self.flat, self.cfg = K, flat, cfg
        dt = torch.float32 if precision == "fp32" else torch.bfloat16
        boc = cfg.block_out_channels
        common = dict(precision=precision, K=K)
        self.blocks = []            # (kind, trainer) in forward order
        h, w, cin = H, W, boc[0]
        for i, cout in enumerate(boc):
            for j in range(cfg.layers_per_block):
                self.blocks.append(("resnet", ResnetBlockTrainer(flat, f"down_blocks.{i}.resnets.{j}", B=B, H=h, W=w, Cin=cin, Cout=cout,
                                                                 groups=cfg.norm_num_groups, eps=cfg.norm_eps, **common)))
                cin = cout
            if i != len(boc) - 1:
                self.blocks.append(("down", DownsampleTrainer(flat, f"down_blocks.{i}.downsamplers.0", B=B, H=h, W=w, C=cout, **common)))
                h, w = h // 2, w // 2
        self.mid = [ResnetBlockTrainer(flat, f"mid_block.resnets.{j}", B=B, H=h, W=w, Cin=boc[-1], Cout=boc[-1], groups=cfg.norm_num_groups,
                                       eps=cfg.norm_eps, **common) for j in range(2)]
        # hidden states in forward order: h0 (the input), then each block's output; consumer k of hidden k is block k (or mid 0)
        self.h0 = torch.zeros(B, H * W, boc[0], device=flat.param.device, dtype=dt)
        hidden = [(self.h0, H, W, boc[0])]
        hh, ww = H, W
        for kind, t in self.blocks:
            if kind == "down":
                hh, ww = hh // 2, ww // 2
            hidden.append((t.out, hh, ww, t.out.shape[-1]))
        consumers = [t for _, t in self.blocks] + [self.mid[0]]
        self.hidden, self._mid_hw = hidden, (h, w)
        skip_grads, mid_other = self._build_up_path(B, dt, common)          # (None, None) here; the full branch overrides it
        self.taps = [ZeroConvTap(flat, f"brushnet_down_blocks.{k}", hb, consumers[k].dx, B=B, H=hh_, W=ww_, C=c, dt=dt, K=K,
                                 d_other2=None if skip_grads is None else skip_grads[k])
                     for k, (hb, hh_, ww_, c) in enumerate(hidden)]
        self.mid_tap = ZeroConvTap(flat, "brushnet_mid_block", self.mid[1].out, mid_other, B=B, H=h, W=w, C=boc[-1], dt=dt, K=K)
        # No copies between blocks: a block reads its producer's output buffer and the gradient buffer of its output's zero-conv
        # (which already holds the sum over all consumers) in place.
        x = self.h0
        for k, (_, t) in enumerate(self.blocks):
            t.rebind(x=x, d_out=self.taps[k + 1].dh)
            x = t.out
        self.mid[0].rebind(x=x, d_out=self.mid[1].dx)
        self.mid[1].rebind(x=self.mid[0].out, d_out=self.mid_tap.dh)

    def _build_up_path(self, B, dt, common):
        return None, None

    def resnet_prefixes(self):
        return [t.p for kind, t in self.blocks if kind == "resnet"] + [t.p for t in self.mid]

    def refresh_dgrad_weights(self):
        for _, t in self.blocks:
            t.refresh_dgrad_weights()
        for t in self.mid + self.taps + [self.mid_tap]:
            t.refresh_dgrad_weights()

    def forward(self, h0: torch.Tensor, rowbias: Dict[str, torch.Tensor]):
        """h0: conv_in_condition's output [B, HW, C0]; rowbias[prefix] = time_emb_proj(silu(emb)) of each resnet ([B, Cout] fp32).
        Returns (down taps [12 for SD1.5], mid tap) — buffers owned by the trainer."""
        if h0.data_ptr() != self.h0.data_ptr():
            self.h0.copy_(h0.view_as(self.h0))
        x = self.h0
        for kind, t in self.blocks:
            x = t.forward(x, rowbias[t.p]) if kind == "resnet" else t.forward(x)
        for t in self.mid:
            x = t.forward(x, rowbias[t.p])
        return [z.forward() for z in self.taps], self.mid_tap.forward()

    def backward(self, d_down_taps, d_mid_tap):
        """Gradients of the 12 + 1 taps in.  Accumulates every parameter gradient into the flat buffer; returns
        (d h0, {resnet prefix: d rowbias})."""
        d_rb = {}
        d = self.mid_tap.backward(d_mid_tap)                       # gradient at mid.resnets.1's output (its only consumer)
        for t in reversed(self.mid):
            d, d_rb[t.p] = t.backward(d)                           # leaves mid[0].dx = gradient from the mid block at the last hidden state
        for k in range(len(self.blocks), 0, -1):                   # hidden k is produced by block k-1, consumed by block k (or mid[0])
            d = self.taps[k].backward(d_down_taps[k])              # W_k^T d tap_k + consumer's dx (residual epilogue)
            kind, t = self.blocks[k - 1]
            if kind == "resnet":
                _, d_rb[t.p] = t.backward(d)
            else:
                t.backward(d)
        return self.taps[0].backward(d_down_taps[0]), d_rb


class UpsampleTrainer:
    """Upsample2D (nearest x2 + conv3x3; S/models/upsampling.py:145-186) forward and backward on the kernels.
    Forward: the sub-pixel `up2x` plan (no upsampled tensor).  Backward: the weight gradient needs the upsampled input once
    (mfb_upsample2x + the stride-1 weight-gradient kernel); the data gradient is the stride-1 data-gradient plan at high
    resolution followed by a 2x2 sum-pool, which runs as a stride-2 plan with a fixed 0/1 weight (taps kh, kw in {1, 2},
    identity over channels) — a first version on existing kernels; a pooling kernel would do it in 4 adds per element.
    Flat entries: `<prefix>.conv.weight` packed [C, 9*C], `<prefix>.conv.bias` [C].
    STATUS: verified on the CPU stand-in (tests/test_oracle_train.py) and, as part of BrushNetTrainer, on B200 in fp32 parity mode."""

    def __init__(self, flat: FlatParams, prefix: str, *, B: int, H: int, W: int, C: int, precision: str = "bf16", K=None):
        K = _ops if K is None else K
        self.K, self.flat, self.p, self.B, self.H, self.W, self.C = K, flat, prefix, B, H, W, C
        self.dt = torch.float32 if precision == "fp32" else torch.bfloat16
        dev = flat.param.device
        self._wsrc = flat.p if self.dt == torch.float32 else flat.w
        z = lambda hw: torch.zeros(B, hw, C, device=dev, dtype=self.dt)
        self.x, self.dx = z(H * W), z(H * W)
        self.out, self.d_out, self.xu, self.du = z(4 * H * W), z(4 * H * W), z(4 * H * W), z(4 * H * W)
        self.w_up = torch.zeros(4, C, 4 * C, device=dev, dtype=self.dt)
        self.wd = torch.zeros(C, 9 * C, device=dev, dtype=self.dt)
        pool = torch.zeros(C, 3, 3, C, device=dev, dtype=torch.float32)
        eye = torch.eye(C, device=dev)
        for kh in (1, 2):
            for kw in (1, 2):
                pool[:, kh, kw, :] = eye
        self.w_pool = pool.reshape(C, 9 * C).to(self.dt).contiguous()
        self._make_plans()
        self.refresh_dgrad_weights()

    def _make_plans(self):
        K, B, H, W, C = self.K, self.B, self.H, self.W, self.C
        self.plan = K.ConvPlan(self.x, self.w_up, self.out, B=B, H=H, W=W, Cin=C, Cout=C, ksize=3, up2x=True, bias=self.flat.p(f"{self.p}.conv.bias"))
        self.plan_d = K.ConvPlan(self.d_out, self.wd, self.du, B=B, H=2 * H, W=2 * W, Cin=C, Cout=C, ksize=3)
        self.plan_pool = K.ConvPlan(self.du, self.w_pool, self.dx, B=B, H=2 * H, W=2 * W, Cin=C, Cout=C, ksize=3, stride=2)

    def rebind(self, x=None, d_out=None):
        if x is not None:
            self.x = x.view_as(self.x)
        if d_out is not None:
            self.d_out = d_out.view_as(self.d_out)
        self._make_plans()

    def refresh_dgrad_weights(self):
        """Derived weight copies: the four sub-pixel phases of the forward plan and the flipped / transposed data-gradient weight."""
        wp = self._wsrc(f"{self.p}.conv.weight")
        old = _ops._ACT[0]
        _ops._ACT[0] = self.dt                     # pack_upconv_weight rounds to the activation dtype of the engine being built
        try:
            self.w_up.copy_(_ops.pack_upconv_weight(unpack_conv_grad(wp.float(), 3)))
        finally:
            _ops._ACT[0] = old
        self.K.dgrad_repack(wp, self.wd, 3)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x.view_as(self.x))
        self.plan.run()
        return self.out

    def backward(self, d_out: torch.Tensor) -> torch.Tensor:
        K, f, B, H, W = self.K, self.flat, self.B, self.H, self.W
        if d_out.data_ptr() != self.d_out.data_ptr():
            self.d_out.copy_(d_out.view_as(self.d_out))
        # nearest x2 of the saved input: a pure pixel-vector copy, so fp32 tensors go through the bf16 kernel as 2*C "channels"
        bf = torch.bfloat16
        K.upsample2x(self.x.view(bf) if self.dt != bf else self.x, self.xu.view(bf) if self.dt != bf else self.xu, B=B, H=H, W=W)
        K.conv_wgrad(self.xu, self.d_out, f.g(f"{self.p}.conv.weight"), f.g(f"{self.p}.conv.bias"), B=B, H=2 * H, W=2 * W, ksize=3,
                     accumulate=True)
        self.plan_d.run()            # gradient at the upsampled tensor
        self.plan_pool.run()         # sum over each 2x2 block = gradient at the low-resolution input
        return self.dx


def skip_resnet_param_shapes(prefix: str, C1: int, C2: int, Cout: int) -> Dict[str, Tuple[int, ...]]:
    """Flat entries of an up-block resnet over cat([x, skip]) (S/models/unets/unet_2d_blocks.py:2711-2728).  The 1x1 shortcut
    weight is stored as its two column blocks (`.a` over x, `.b` over the skip) so each half is a dense matrix for the
    weight-gradient kernel and for its own data-gradient plan."""
    s = resnet_param_shapes(prefix, C1 + C2, Cout)
    del s[f"{prefix}.conv_shortcut.weight"]
    s[f"{prefix}.conv_shortcut.weight.a"] = (Cout, C1)
    s[f"{prefix}.conv_shortcut.weight.b"] = (Cout, C2)
    return s


def pack_skip_resnet_state_dict(prefix: str, sd: Dict[str, torch.Tensor], C1: int) -> Dict[str, torch.Tensor]:
    out = pack_resnet_state_dict(prefix, sd)
    w = out.pop(f"{prefix}.conv_shortcut.weight")
    out[f"{prefix}.conv_shortcut.weight.a"] = w[:, :C1].contiguous()
    out[f"{prefix}.conv_shortcut.weight.b"] = w[:, C1:].contiguous()
    return out


class SkipResnetBlockTrainer:
    """ResnetBlock2D over the channel concat of two tensors (x from below, skip from the down path) — the resnets of
    BrushNet's up blocks.  The concat is never materialised except normalised (GroupNorm reads both sources, as in the
    inference engine).  Backward returns one gradient per source: the GroupNorm backward writes its two halves, and the two
    data-gradient plans of the shortcut halves add them in their residual epilogue (d x = W_a^T d out + GN-half; same for
    the skip).  STATUS: verified on the CPU stand-in and, as part of BrushNetTrainer, on B200 in fp32 parity mode."""

    def __init__(self, flat: FlatParams, prefix: str, *, B: int, H: int, W: int, C1: int, C2: int, Cout: int, groups: int = 32,
                 eps: float = 1e-5, precision: str = "bf16", K=None):
        K = _ops if K is None else K
        self.K, self.flat, self.p = K, flat, prefix
        self.B, self.H, self.W, self.HW, self.C1, self.C2, self.Cout, self.groups, self.eps = B, H, W, H * W, C1, C2, Cout, groups, eps
        Cin = C1 + C2
        self.dt = torch.float32 if precision == "fp32" else torch.bfloat16
        dev = flat.param.device
        act = lambda c: torch.zeros(B, self.HW, c, device=dev, dtype=self.dt)
        self.x1, self.x2, self.n1, self.c1, self.n2 = act(C1), act(C2), act(Cin), act(Cout), act(Cout)
        self.sc_a, self.sc, self.out = act(Cout), act(Cout), act(Cout)
        self.rowbias = torch.zeros(B, Cout, device=dev, dtype=torch.float32)
        self.d_out, self.dn2, self.dc1, self.dn1 = act(Cout), act(Cout), act(Cout), act(Cin)
        self.g1, self.g2, self.dx, self.dx2 = act(C1), act(C2), act(C1), act(C2)     # self.dx: gradient of the main input (like the other trainers)
        self.d_rowbias = torch.zeros(B, Cout, device=dev, dtype=torch.float32)
        # one GroupNorm workspace per norm: the forward leaves its per-(image, group) {sum, sum of squares} at the head of it, and
        # the backward reads them there instead of re-reading the tensor for a statistics pass (maps above 8x8: the single-launch
        # kernel of the small maps keeps its statistics in registers)
        self.gn_ws = torch.zeros(K.gn_ws_floats(B, groups), device=dev, dtype=torch.float32)
        self.gn_ws2 = torch.zeros(K.gn_ws_floats(B, groups), device=dev, dtype=torch.float32)
        self.keep_stats = self.HW > 64 and os.environ.get("MFB_TRAIN_KEEP_GN_STATS", "1") == "1"      # =0: recompute in the backward (A/B)
        self.gnb_ws = torch.zeros(2 * B * max(Cin, Cout), device=dev, dtype=torch.float32)
        self._wsrc = wsrc = flat.p if self.dt == torch.float32 else flat.w
        n = lambda s: f"{prefix}.{s}"
        geo = dict(B=B, H=H, W=W)
        self.w1d = torch.zeros(Cin, 9 * Cout, device=dev, dtype=self.dt)
        self.w2d = torch.zeros(Cout, 9 * Cout, device=dev, dtype=self.dt)
        self.wad = torch.zeros(C1, Cout, device=dev, dtype=self.dt)
        self.wbd = torch.zeros(C2, Cout, device=dev, dtype=self.dt)
        self._make_plans()
        self.refresh_dgrad_weights()

    def _make_plans(self):
        K, flat, wsrc = self.K, self.flat, self._wsrc
        C1, C2, Cout = self.C1, self.C2, self.Cout
        Cin = C1 + C2
        n = lambda s: f"{self.p}.{s}"
        geo = dict(B=self.B, H=self.H, W=self.W)
        self.plan1 = K.ConvPlan(self.n1, wsrc(n("conv1.weight")), self.c1, Cin=Cin, Cout=Cout, ksize=3, bias=flat.p(n("conv1.bias")),
                                rowbias=self.rowbias, rowbias_ld=Cout, **geo)
        self.plan_sc_a = K.ConvPlan(self.x1, wsrc(n("conv_shortcut.weight.a")), self.sc_a, Cin=C1, Cout=Cout, ksize=1,
                                    bias=flat.p(n("conv_shortcut.bias")), **geo)
        self.plan_sc_b = K.ConvPlan(self.x2, wsrc(n("conv_shortcut.weight.b")), self.sc, Cin=C2, Cout=Cout, ksize=1, res1=self.sc_a, **geo)
        self.plan2 = K.ConvPlan(self.n2, wsrc(n("conv2.weight")), self.out, Cin=Cout, Cout=Cout, ksize=3, bias=flat.p(n("conv2.bias")),
                                res1=self.sc, **geo)
        self.plan_d2 = K.ConvPlan(self.d_out, self.w2d, self.dn2, Cin=Cout, Cout=Cout, ksize=3, **geo)
        self.plan_d1 = K.ConvPlan(self.dc1, self.w1d, self.dn1, Cin=Cout, Cout=Cin, ksize=3, **geo)
        self.plan_da = K.ConvPlan(self.d_out, self.wad, self.dx, Cin=Cout, Cout=C1, ksize=1, res1=self.g1, **geo)
        self.plan_db = K.ConvPlan(self.d_out, self.wbd, self.dx2, Cin=Cout, Cout=C2, ksize=1, res1=self.g2, **geo)

    def rebind(self, x1=None, x2=None, d_out=None):
        if x1 is not None:
            self.x1 = x1.view_as(self.x1)
        if x2 is not None:
            self.x2 = x2.view_as(self.x2)
        if d_out is not None:
            self.d_out = d_out.view_as(self.d_out)
        self._make_plans()

    def refresh_dgrad_weights(self):
        n = lambda s: f"{self.p}.{s}"
        R = self.K.dgrad_repack
        R(self._wsrc(n("conv1.weight")), self.w1d, 3)
        R(self._wsrc(n("conv2.weight")), self.w2d, 3)
        R(self._wsrc(n("conv_shortcut.weight.a")), self.wad, 1)
        R(self._wsrc(n("conv_shortcut.weight.b")), self.wbd, 1)

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, rowbias: Optional[torch.Tensor] = None) -> torch.Tensor:
        K, f, n = self.K, self.flat, (lambda s: f"{self.p}.{s}")
        if x1.data_ptr() != self.x1.data_ptr():
            self.x1.copy_(x1.view_as(self.x1))
        if x2.data_ptr() != self.x2.data_ptr():
            self.x2.copy_(x2.view_as(self.x2))
        if rowbias is None:
            self.rowbias.zero_()
        else:
            self.rowbias.copy_(rowbias)
        gn = dict(B=self.B, HW=self.HW, groups=self.groups, eps=self.eps, silu=True)
        K.groupnorm(self.x1, self.x2, f.p(n("norm1.weight")), f.p(n("norm1.bias")), self.n1, self.gn_ws, **gn)
        self.plan1.run()
        K.groupnorm(self.c1, None, f.p(n("norm2.weight")), f.p(n("norm2.bias")), self.n2, self.gn_ws2, **gn)
        self.plan_sc_a.run()
        self.plan_sc_b.run()
        self.plan2.run()
        return self.out

    def backward(self, d_out: torch.Tensor):
        """-> (d x, d skip, d rowbias)."""
        K, f, n = self.K, self.flat, (lambda s: f"{self.p}.{s}")
        B, H, W, HW = self.B, self.H, self.W, self.HW
        if d_out.data_ptr() != self.d_out.data_ptr():
            self.d_out.copy_(d_out.view_as(self.d_out))
        gnb = dict(B=B, HW=HW, groups=self.groups, eps=self.eps, silu=True, accumulate=True)
        wg = dict(B=B, H=H, W=W, accumulate=True)
        K.conv_wgrad(self.n2, self.d_out, f.g(n("conv2.weight")), f.g(n("conv2.bias")), ksize=3, **wg)
        self.plan_d2.run()
        K.groupnorm_bwd(self.c1, None, self.dn2, f.p(n("norm2.weight")), f.p(n("norm2.bias")), self.dc1, None, self.gnb_ws,
                        dgamma=f.g(n("norm2.weight")), dbeta=f.g(n("norm2.bias")), stats=self.gn_ws2 if self.keep_stats else None, **gnb)
        K.conv_wgrad(self.n1, self.dc1, f.g(n("conv1.weight")), f.g(n("conv1.bias")), ksize=3, **wg)
        K.rowsum_per_image(self.dc1, self.d_rowbias, B=B, HW=HW)
        self.plan_d1.run()
        K.groupnorm_bwd(self.x1, self.x2, self.dn1, f.p(n("norm1.weight")), f.p(n("norm1.bias")), self.g1, self.g2, self.gnb_ws,
                        dgamma=f.g(n("norm1.weight")), dbeta=f.g(n("norm1.bias")), stats=self.gn_ws if self.keep_stats else None, **gnb)
        K.conv_wgrad(self.x1, self.d_out, f.g(n("conv_shortcut.weight.a")), f.g(n("conv_shortcut.bias")), ksize=1, **wg)
        K.conv_wgrad(self.x2, self.d_out, f.g(n("conv_shortcut.weight.b")), None, ksize=1, **wg)
        self.plan_da.run()            # d x    = W_a^T d out + GroupNorm-backward half 1
        self.plan_db.run()            # d skip = W_b^T d out + GroupNorm-backward half 2
        return self.dx, self.dx2, self.d_rowbias


def brushnet_branch_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """Flat entries of the whole BrushNet branch behind conv_in_condition: down path, mid block, up blocks and all 28 zero-convs
    (SD1.5: 12 + 1 + 15; S/models/brushnet.py:296-449)."""
    s = brushnet_down_mid_shapes(cfg)
    boc = cfg.block_out_channels
    down_c = [boc[0]]
    for i, c in enumerate(boc):
        down_c += [c] * cfg.layers_per_block + ([c] if i != len(boc) - 1 else [])
    rev = list(reversed(boc))
    cx, k, nl = boc[-1], 0, cfg.layers_per_block + 1
    for i, cout in enumerate(rev):
        for j in range(nl):
            cskip = down_c[len(down_c) - 1 - (i * nl + j)]
            s.update(skip_resnet_param_shapes(f"up_blocks.{i}.resnets.{j}", cx, cskip, cout))
            cx = cout
            s[f"brushnet_up_blocks.{k}.weight"], s[f"brushnet_up_blocks.{k}.bias"] = (cout, cout), (cout,)
            k += 1
        if i != len(rev) - 1:
            s[f"up_blocks.{i}.upsamplers.0.conv.weight"], s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (cout, 9 * cout), (cout,)
            s[f"brushnet_up_blocks.{k}.weight"], s[f"brushnet_up_blocks.{k}.bias"] = (cout, cout), (cout,)
            k += 1
    return s


def pack_brushnet_branch(cfg, sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference state_dict (OIHW) -> fp32 tensors in the layout of brushnet_branch_shapes."""
    out = {}
    for name, shp in brushnet_branch_shapes(cfg).items():
        if name.endswith(".conv_shortcut.weight.a") or name.endswith(".conv_shortcut.weight.b"):
            w = sd[name[:-2]].float()[:, :, 0, 0]
            c1 = brushnet_branch_shapes(cfg)[name[:-2] + ".a"][1]
            out[name] = (w[:, :c1] if name.endswith(".a") else w[:, c1:]).contiguous()
            continue
        v = sd[name].float()
        out[name] = (v.permute(0, 2, 3, 1).reshape(v.shape[0], -1) if v.dim() == 4 else v).reshape(shp).contiguous()
    return out


class BrushNetBranchTrainer(BrushNetDownMidTrainer):
    """The whole BrushNet branch behind conv_in_condition, forward-for-training and backward: down path, MidBlock2D, the up
    blocks (resnets over cat([x, skip]), Upsample2D) and all 28 zero-conv taps (oracle/mf_oracle.py `brushnet_forward`;
    S/models/brushnet.py:810-906).  A down-path hidden state now has up to three consumers — the next block, its zero-conv and
    the up-block resnet that takes it as skip — and its gradient is still formed inside ONE data-gradient plan (the zero-conv's,
    with both other gradients as residual inputs).  Backward order: up path in reverse (which leaves every skip gradient in the
    dx2 buffer of its resnet), mid block, down path.
    STATUS: verified on the CPU stand-in (tests/test_oracle_train.py) and on B200 in fp32 parity mode (down / mid half on its own,
    the whole branch as part of BrushNetTrainer: tests/test_gpu_zz_train_net.py)."""

    def _build_up_path(self, B, dt, common):
        cfg, flat = self.cfg, self.flat
        boc = cfg.block_out_channels
        rev = list(reversed(boc))
        nl = cfg.layers_per_block + 1
        h, w = self._mid_hw
        n_hidden = len(self.hidden)
        self.up_seq = []                # (kind, trainer, skip hidden index or None) in forward order
        cx = boc[-1]
        for i, cout in enumerate(rev):
            for j in range(nl):
                idx = n_hidden - 1 - (i * nl + j)
                cskip = self.hidden[idx][3]
                self.up_seq.append(("resnet", SkipResnetBlockTrainer(flat, f"up_blocks.{i}.resnets.{j}", B=B, H=h, W=w, C1=cx, C2=cskip, Cout=cout,
                                                                     groups=cfg.norm_num_groups, eps=cfg.norm_eps, **common), idx))
                cx = cout
            if i != len(rev) - 1:
                self.up_seq.append(("up", UpsampleTrainer(flat, f"up_blocks.{i}.upsamplers.0", B=B, H=h, W=w, C=cout, **common), None))
                h, w = 2 * h, 2 * w
        # zero-convs of the up hidden states: consumer = the next block of the up sequence (none for the last)
        self.up_taps = []
        hh, ww = self._mid_hw
        for k, (kind, t, _) in enumerate(self.up_seq):
            if kind == "up":
                hh, ww = 2 * hh, 2 * ww
            nxt = self.up_seq[k + 1][1].dx if k + 1 < len(self.up_seq) else None
            self.up_taps.append(ZeroConvTap(flat, f"brushnet_up_blocks.{k}", t.out, nxt, B=B, H=hh, W=ww, C=t.out.shape[-1], dt=dt, K=self.K))
        skip_grads = [None] * n_hidden
        for kind, t, idx in self.up_seq:
            if kind == "resnet":
                skip_grads[idx] = t.dx2
        x = self.mid[1].out                    # no copies between blocks (see BrushNetDownMidTrainer.__init__)
        for k, (kind, t, idx) in enumerate(self.up_seq):
            if kind == "resnet":
                t.rebind(x1=x, x2=self.hidden[idx][0], d_out=self.up_taps[k].dh)
            else:
                t.rebind(x=x, d_out=self.up_taps[k].dh)
            x = t.out
        return skip_grads, self.up_seq[0][1].dx

    def resnet_prefixes(self):
        return super().resnet_prefixes() + [t.p for kind, t, _ in self.up_seq if kind == "resnet"]

    def refresh_dgrad_weights(self):
        super().refresh_dgrad_weights()
        for _, t, _ in self.up_seq:
            t.refresh_dgrad_weights()
        for t in self.up_taps:
            t.refresh_dgrad_weights()

    def forward(self, h0: torch.Tensor, rowbias: Dict[str, torch.Tensor]):
        """-> (down taps, mid tap, up taps)."""
        down_taps, mid_tap = super().forward(h0, rowbias)
        x = self.mid[1].out
        for kind, t, idx in self.up_seq:
            x = t.forward(x, self.hidden[idx][0], rowbias[t.p]) if kind == "resnet" else t.forward(x)
        return down_taps, mid_tap, [z.forward() for z in self.up_taps]

    after_up_backward = None      # optional hook: called once the up path's parameter gradients are complete (overlapped all-reduce)

    def backward(self, d_down_taps, d_mid_tap, d_up_taps):
        d_rb = {}
        for k in range(len(self.up_seq) - 1, -1, -1):
            d = self.up_taps[k].backward(d_up_taps[k])             # + the next up block's dx
            kind, t, _ = self.up_seq[k]
            if kind == "resnet":
                _, _, d_rb[t.p] = t.backward(d)                    # leaves d x in t.dx and the skip gradient in t.dx2
            else:
                t.backward(d)
        if self.after_up_backward is not None:
            self.after_up_backward()
        d_h0, d_rb_down = super().backward(d_down_taps, d_mid_tap)  # mid tap adds up_seq[0].dx; down taps add the skip gradients
        d_rb.update(d_rb_down)
        return d_h0, d_rb


def unpack_brushnet_branch(cfg, flat: FlatParams) -> Dict[str, torch.Tensor]:
    """Inverse of pack_brushnet_branch: the trained branch parameters back in the reference's state_dict naming and OIHW layout
    (what `BrushNetModel.save_pretrained` writes at a checkpoint, E/train_brushnet_mirror.py:997-1032) — fp32 masters, on the host.
    Only the branch behind conv_in_condition; `unpack_brushnet` exports the whole model."""
    shapes = brushnet_branch_shapes(cfg)
    out: Dict[str, torch.Tensor] = {}
    for name, shp in shapes.items():
        if name.endswith(".conv_shortcut.weight.b"):
            continue
        v = flat.p(name).detach().float().cpu()
        if name.endswith(".conv_shortcut.weight.a"):
            v = torch.cat([v, flat.p(name[:-2] + ".b").detach().float().cpu()], 1)
            out[name[:-2]] = v[:, :, None, None].contiguous()
        elif name.endswith(".weight") and len(shp) == 2 and ".norm" not in name:
            Cout, k = shp
            if name.startswith("brushnet_") or ".conv_shortcut." in name:      # 1x1 convs
                out[name] = v[:, :, None, None].contiguous()
            else:                                                            # packed 3x3: [Cout, (kh, kw, ci)] -> OIHW
                out[name] = unpack_conv_grad(v, 3)
        else:
            out[name] = v.clone()
    return out


def time_path_shapes(cfg, resnet_prefixes) -> Dict[str, Tuple[int, ...]]:
    """Flat entries of the timestep path: the TimestepEmbedding MLP and ALL time_emb_proj layers of the net concatenated into
    one [sum Cout, temb] matrix in the order of `resnet_prefixes` (one GEMV forward, one weight-gradient call backward)."""
    c0, temb = cfg.block_out_channels[0], cfg.time_embed_dim
    return {"time_embedding.linear_1.weight": (temb, c0), "time_embedding.linear_1.bias": (temb,),
            "time_embedding.linear_2.weight": (temb, temb), "time_embedding.linear_2.bias": (temb,),
            "time_emb_proj.wcat": (sum(resnet_cout(cfg, p) for p in resnet_prefixes), temb),
            "time_emb_proj.bcat": (sum(resnet_cout(cfg, p) for p in resnet_prefixes),)}


def resnet_cout(cfg, prefix: str) -> int:
    """Output channels of a BrushNet resnet by its state_dict prefix."""
    boc = cfg.block_out_channels
    parts = prefix.split(".")
    if parts[0] == "mid_block":
        return boc[-1]
    i = int(parts[1])
    return boc[i] if parts[0] == "down_blocks" else list(reversed(boc))[i]


def pack_time_path(cfg, sd, resnet_prefixes) -> Dict[str, torch.Tensor]:
    out = {k: sd[k].float().contiguous() for k in ("time_embedding.linear_1.weight", "time_embedding.linear_1.bias",
                                                    "time_embedding.linear_2.weight", "time_embedding.linear_2.bias")}
    out["time_emb_proj.wcat"] = torch.cat([sd[p + ".time_emb_proj.weight"].float() for p in resnet_prefixes], 0).contiguous()
    out["time_emb_proj.bcat"] = torch.cat([sd[p + ".time_emb_proj.bias"].float() for p in resnet_prefixes], 0).contiguous()
    return out


class TimePathTrainer:
    """Timestep path forward and backward (S/models/embeddings.py:27-67,226-237; S/models/resnet.py:369-376):
    sinusoid(t) -> linear_1 -> SiLU -> linear_2 -> SiLU -> the concatenated time_emb_proj GEMV -> row-bias table [B, sum Cout].
    Backward consumes the per-resnet d rowbias of the block programs.  Everything is GEMV-sized (M = batch): the weight gradients
    are `conv_wgrad` calls with the batch as the "pixels", the data gradients are one implicit-GEMM plan (K = sum Cout is too long
    for the small-linear kernel) and one small linear, the activation derivative is `silu_bwd`.
    STATUS: verified on the CPU stand-in and, as part of BrushNetTrainer, on B200 in fp32 parity mode."""

    def __init__(self, flat: FlatParams, cfg, resnet_prefixes, *, B: int, precision: str = "bf16", K=None):
        K = _ops if K is None else K
        self.K, self.flat, self.cfg, self.B, self.prefixes = K, flat, cfg, B, list(resnet_prefixes)
        self.dt = torch.float32 if precision == "fp32" else torch.bfloat16
        dev = flat.param.device
        c0, temb = cfg.block_out_channels[0], cfg.time_embed_dim
        self.N = N = sum(resnet_cout(cfg, p) for p in self.prefixes)
        self.off, o = {}, 0
        for p in self.prefixes:
            self.off[p] = (o, resnet_cout(cfg, p))
            o += resnet_cout(cfg, p)
        f32 = torch.float32
        z = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)
        self.t, self.sin, self.h1, self.s1, self.e, self.s2 = z(B), z(B, c0), z(B, temb), z(B, temb), z(B, temb), z(B, temb)
        self.rowbias, self.d_rb = z(B, N), z(B, N)
        self.d_rb_act, self.d_s2_act = z(B, N, dt=self.dt), z(B, temb, dt=self.dt)
        self.d_e, self.d_s1, self.d_h1 = z(B, temb), z(B, temb), z(B, temb)
        self.zero_bias = z(temb)
        # working copies the small-linear kernel reads (activation dtype), and the transposed ones of the data gradients
        self.w1, self.w2, self.wcat = z(temb, c0, dt=self.dt), z(temb, temb, dt=self.dt), z(N, temb, dt=self.dt)
        self.w2t, self.wcat_t = z(temb, temb, dt=self.dt), z(temb, N, dt=self.dt)
        self.plan_ds2 = K.ConvPlan(self.d_rb_act, self.wcat_t, self.d_s2_act, B=1, H=1, W=B, Cin=N, Cout=temb, ksize=1)
        self.refresh_dgrad_weights()

    def refresh_dgrad_weights(self):
        f = self.flat
        self.w1.copy_(f.p("time_embedding.linear_1.weight"))
        self.w2.copy_(f.p("time_embedding.linear_2.weight"))
        self.wcat.copy_(f.p("time_emb_proj.wcat"))
        self.w2t.copy_(f.p("time_embedding.linear_2.weight").t())
        self.wcat_t.copy_(f.p("time_emb_proj.wcat").t())

    def rowbias_of(self, prefix: str) -> torch.Tensor:
        o, c = self.off[prefix]
        return self.rowbias[:, o:o + c]

    def forward(self, timesteps: torch.Tensor) -> Dict[str, torch.Tensor]:
        """timesteps [B] (any numeric dtype) -> {resnet prefix: rowbias view [B, Cout]}."""
        K, f = self.K, self.flat
        self.t.copy_(timesteps.to(torch.float32))
        K.timestep_sinusoid(self.t, self.sin)
        K.linear_small(self.sin, self.w1, f.p("time_embedding.linear_1.bias"), self.h1)                 # pre-activation kept
        K.linear_small(self.h1, self.w2, f.p("time_embedding.linear_2.bias"), self.e, act_in=True)
        K.linear_small(self.e, self.wcat, f.p("time_emb_proj.bcat"), self.rowbias, act_in=True)
        return {p: self.rowbias_of(p) for p in self.prefixes}

    def backward(self, d_rowbias: Dict[str, torch.Tensor]):
        K, f, B = self.K, self.flat, self.B
        for p, g in d_rowbias.items():
            o, c = self.off[p]
            self.d_rb[:, o:o + c].copy_(g)
        wg = dict(B=B, H=1, W=1, ksize=1, accumulate=True)
        K.silu_bwd(self.e, y=self.s2)                                                              # silu(e), recomputed
        K.conv_wgrad(self.s2, self.d_rb, f.g("time_emb_proj.wcat"), f.g("time_emb_proj.bcat"), **wg)
        K.f32_to_bf16(self.d_rb, self.d_rb_act)
        self.plan_ds2.run()                                                                        # d silu(e) = d rb . Wcat
        K.silu_bwd(self.e, dy=self.d_s2_act, dx=self.d_e)
        K.silu_bwd(self.h1, y=self.s1)
        K.conv_wgrad(self.s1, self.d_e, f.g("time_embedding.linear_2.weight"), f.g("time_embedding.linear_2.bias"), **wg)
        K.linear_small(self.d_e, self.w2t, self.zero_bias, self.d_s1)                              # d silu(h1) = d e . W2
        K.silu_bwd(self.h1, dy=self.d_s1, dx=self.d_h1)
        K.conv_wgrad(self.sin, self.d_h1, f.g("time_embedding.linear_1.weight"), f.g("time_embedding.linear_1.bias"), **wg)


def brushnet_resnet_prefixes(cfg):
    """state_dict prefixes of BrushNet's 22 resnets in forward order (down, mid, up)."""
    n, nl = len(cfg.block_out_channels), cfg.layers_per_block
    return ([f"down_blocks.{i}.resnets.{j}" for i in range(n) for j in range(nl)] + [f"mid_block.resnets.{j}" for j in range(2)] +
            [f"up_blocks.{i}.resnets.{j}" for i in range(n) for j in range(nl + 1)])


def brushnet_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """Flat entries of the WHOLE BrushNetModel: conv_in_condition (packed), the timestep path, the branch and its 28 zero-convs."""
    cin = cfg.in_channels + cfg.conditioning_channels
    s = {"conv_in_condition.weight": (cfg.block_out_channels[0], 9 * cin), "conv_in_condition.bias": (cfg.block_out_channels[0],)}
    s.update(time_path_shapes(cfg, brushnet_resnet_prefixes(cfg)))
    s.update(brushnet_branch_shapes(cfg))
    return s


def pack_brushnet(cfg, sd) -> Dict[str, torch.Tensor]:
    w = sd["conv_in_condition.weight"].float()
    out = {"conv_in_condition.weight": w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous(),
           "conv_in_condition.bias": sd["conv_in_condition.bias"].float().contiguous()}
    out.update(pack_time_path(cfg, sd, brushnet_resnet_prefixes(cfg)))
    out.update(pack_brushnet_branch(cfg, sd))
    return out


def unpack_brushnet(cfg, flat: FlatParams) -> Dict[str, torch.Tensor]:
    """Inverse of pack_brushnet: EVERY parameter BrushNetTrainer trains, back in the reference's state_dict naming and layouts
    (`BrushNetModel.save_pretrained` at a checkpoint, E/train_brushnet_mirror.py:997-1032): the branch, conv_in_condition (OIHW)
    and the timestep path — the concatenated `time_emb_proj.wcat` / `.bcat` split back into the 22 per-resnet layers.  The key set
    passes `checkpoint.check_state_dict(..., "brushnet")`; fp32 masters, on the host."""
    out = unpack_brushnet_branch(cfg, flat)
    host = lambda name: flat.p(name).detach().float().cpu()
    out["conv_in_condition.weight"] = unpack_conv_grad(host("conv_in_condition.weight"), 3)
    out["conv_in_condition.bias"] = host("conv_in_condition.bias").clone()
    for k in ("time_embedding.linear_1.weight", "time_embedding.linear_1.bias", "time_embedding.linear_2.weight", "time_embedding.linear_2.bias"):
        out[k] = host(k).clone()
    wcat, bcat, o = host("time_emb_proj.wcat"), host("time_emb_proj.bcat"), 0
    for p in brushnet_resnet_prefixes(cfg):
        c = resnet_cout(cfg, p)
        out[p + ".time_emb_proj.weight"] = wcat[o:o + c].clone()
        out[p + ".time_emb_proj.bias"] = bcat[o:o + c].clone()
        o += c
    return out


class BrushNetTrainer:
    """BrushNetModel.forward (S/models/brushnet.py:678-925) for training and its backward, end to end on the kernels:
    (sample, brushnet_cond, timesteps) -> 12 + 1 + 15 taps; backward(d taps) accumulates the gradient of EVERY BrushNet parameter
    (618.8 M for SD1.5) into the flat buffer.  = conv_in_condition (boundary kernel forward, CUDA-core weight gradient over its
    10 input channels) + TimePathTrainer + BrushNetBranchTrainer.  The tap gradients are what the frozen UNet's backward will
    deliver (not built).  conditioning_scale = 1 (training).
    STATUS: verified against autograd of the oracle BrushNet over all parameters on the CPU stand-in AND on B200 in fp32 parity
    mode (tests/test_gpu_zz_train_net.py); bf16 mode not yet run."""

    def __init__(self, flat: FlatParams, cfg, *, B: int, H: int, W: int, precision: str = "bf16", K=None):
        K = _ops if K is None else K
        self.K, self.flat, self.cfg, self.B, self.H, self.W = K, flat, cfg, B, H, W
        self.dt = torch.float32 if precision == "fp32" else torch.bfloat16
        dev = flat.param.device
        self.cin = cfg.in_channels + cfg.conditioning_channels
        c0 = cfg.block_out_channels[0]
        # bf16 product path: the 10-channel input is zero-padded to 64 channels so that conv_in_condition's weight gradient runs on
        # the tcgen05 kernel (the CUDA-core kernel for ragged channel counts took 7.6 ms per step at batch 8, 10 % of the step)
        self.cpad = 64 if (self.dt == torch.bfloat16 and K is _ops and self.cin < 64) else self.cin
        self.xcat = torch.zeros(B, self.cpad, H, W, device=dev, dtype=torch.float32)
        self.xcat_nhwc = torch.zeros(B, H * W, self.cpad, device=dev, dtype=self.dt)
        self.dw_in_pad = torch.zeros(c0, 9 * self.cpad, device=dev, dtype=torch.float32) if self.cpad != self.cin else None
        self.h0 = torch.zeros(B, H * W, c0, device=dev, dtype=self.dt)
        self.w_in = torch.zeros(3, 3, self.cin, c0, device=dev, dtype=torch.float32)
        self.time = TimePathTrainer(flat, cfg, brushnet_resnet_prefixes(cfg), B=B, precision=precision, K=K)
        self.branch = BrushNetBranchTrainer(flat, cfg, B=B, H=H, W=W, precision=precision, K=K)
        self.h0 = self.branch.h0               # conv_in_condition writes straight into the branch's input buffer
        self.refresh_dgrad_weights(parts=False)

    def bind_tap_gradients(self, d_taps):
        """d_taps: the 28 tap-gradient buffers of the frozen UNet's backward (FrozenUNetTrainer.d_taps, pop order): every zero-conv
        then reads its gradient in place instead of a per-step copy."""
        br = self.branch
        zs = list(br.taps) + [br.mid_tap] + list(br.up_taps)
        if len(zs) != len(d_taps):
            raise ValueError(f"expected {len(zs)} tap gradients, got {len(d_taps)}")
        for z, g in zip(zs, d_taps):
            z.bind_d_tap(g)

    def refresh_dgrad_weights(self, parts: bool = True):
        """After every optimizer step: re-derive the weight copies the kernels read in another layout."""
        c0 = self.cfg.block_out_channels[0]
        self.w_in.copy_(self.flat.p("conv_in_condition.weight").view(c0, 3, 3, self.cin).permute(1, 2, 3, 0))
        if parts:
            self.time.refresh_dgrad_weights()
            self.branch.refresh_dgrad_weights()

    def forward(self, sample: torch.Tensor, brushnet_cond: torch.Tensor, timesteps: torch.Tensor):
        """sample [B,4,H,W], brushnet_cond [B,6,H,W] fp32 NCHW (the reference layout), timesteps [B].
        -> (down taps, mid tap, up taps) as NHWC buffers owned by the trainer."""
        K, f = self.K, self.flat
        ca = sample.shape[1]
        self.xcat[:, :ca].copy_(sample)
        self.xcat[:, ca:self.cin].copy_(brushnet_cond)                           # torch.cat([sample, cond], 1), brushnet.py:810
        K.conv_in(self.xcat[:, :ca].contiguous(), self.xcat[:, ca:self.cin].contiguous(), self.w_in, f.p("conv_in_condition.bias"), self.h0)
        rb = self.time.forward(timesteps)
        return self.branch.forward(self.h0, rb)

    def backward(self, d_down_taps, d_mid_tap, d_up_taps):
        K, f = self.K, self.flat
        d_h0, d_rb = self.branch.backward(d_down_taps, d_mid_tap, d_up_taps)
        self.time.backward(d_rb)
        # conv_in_condition: weight / bias gradient over the 10-channel input (NHWC copy in the activation dtype)
        if self.dt == torch.bfloat16:
            K.nchw_to_nhwc(self.xcat, self.xcat_nhwc)
        else:
            self.xcat_nhwc.copy_(self.xcat.permute(0, 2, 3, 1).reshape(self.xcat_nhwc.shape))
        if self.dw_in_pad is None:
            K.conv_wgrad(self.xcat_nhwc, d_h0, f.g("conv_in_condition.weight"), f.g("conv_in_condition.bias"), B=self.B, H=self.H, W=self.W,
                         ksize=3, accumulate=True)
        else:       # padded channels carry zeros: their gradient columns are dropped
            K.conv_wgrad(self.xcat_nhwc, d_h0, self.dw_in_pad, f.g("conv_in_condition.bias"), B=self.B, H=self.H, W=self.W, ksize=3)
            c0 = self.cfg.block_out_channels[0]
            f.g("conv_in_condition.weight").view(c0, 9, self.cin).add_(self.dw_in_pad.view(c0, 9, self.cpad)[:, :, :self.cin])
