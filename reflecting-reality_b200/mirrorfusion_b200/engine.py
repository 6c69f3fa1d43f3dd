"""Host-side sequencing of one denoise step on the libmfb200 kernels.

`StepEngine` owns the two networks of the MirrorFusion hot path for a FIXED problem geometry
(net batch B = 2 x images with CFG, latent H x W): every activation lives in a buffer allocated once, every
GEMM/conv is a prepared plan (TMA descriptors encoded once), so a whole step is a static list of kernel
launches that can be captured into one CUDA graph.  Layer order follows
BrushNetModel.forward (S/models/brushnet.py:678-925) and UNet2DConditionModel.forward
(S/models/unets/unet_2d_condition.py:1039-1348) with the tap sites of SURVEY.md §3.3.

Fusions relative to the reference's op list:
  * skip concat + GroupNorm + SiLU: one kernel reading both sources (no torch.cat tensor);
  * ResnetBlock2D conv2 + 1x1 conv_shortcut (over both concat halves) + identity residual + BrushNet tap:
    one implicit GEMM (shortcut = extra K-segments, residual/tap in the epilogue);
  * conv1 + bias + time_emb_proj(silu(emb)) broadcast: epilogue row-bias, the 22 projections of a net are one GEMV;
  * q/k/v projections: one GEMM (N = 3C); attention out-proj / FF-out / proj_out + residual (+tap): epilogue;
  * GEGLU: epilogue of the FF-in GEMM; zero-conv x conditioning_scale: epilogue alpha;
  * cross-attention K/V of the 77-token context: computed once per prompt, not per step.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .config import NetConfig, tap_channels, up_block_channels

f32 = torch.float32


class _Net:
    def __init__(self, cfg: NetConfig, sd: Dict[str, torch.Tensor], B: int, H: int, W: int, device, name: str, host_pack: bool = False):
        """host_pack: repack the weights (casts, permutes, concatenations, bias sums) where the state_dict lives — normally the
        host — and upload the results, instead of uploading fp32 tensors and repacking with ~600 small device kernels.  Same
        bytes either way; it keeps the device's launch stream to this library's kernels (the driver's launch census of smoke())
        at the price of a slower constructor for the 1.5 G-parameter nets."""
        self.cfg, self.B, self.H, self.W, self.dev, self.name = cfg, B, H, W, device, name
        self.act = ops.act_dtype()     # storage dtype: bf16 (product) or fp32 (parity mode, `with ops.precision('fp32')`)
        if host_pack:
            self.sd = {k: v.detach().to(dtype=f32) for k, v in sd.items()}
        else:
            self.sd = {k: v.detach().to(device=device, dtype=f32) for k, v in sd.items()}
        self.prog: List[Callable[[], None]] = []
        self.tags: List[Tuple[str, float]] = []       # (kernel family, algorithmic FLOPs) per program entry
        self.notes: List[Optional[str]] = []          # shape note per program entry (igemm plans only)
        self.exec_flops: List[float] = []             # FLOPs of the MMAs actually issued per entry (< algorithmic for the sub-pixel upsample)
        self.keep: List[object] = []
        self._scratch: Dict[Tuple, torch.Tensor] = {}
        self.flops = 0.0
        self.launches = 0
        self.n_time_ops = 0
        self.writer_pos: Dict[int, int] = {}      # data_ptr of a produced tensor -> program position of its writer
        self.ext_reads: Dict[int, List[int]] = {}  # program position -> data_ptrs it reads (candidates for cross-net deps)
        # GroupNorm statistics produced by the igemm epilogue that wrote a tensor: data_ptr -> (partials buffer, tiles)
        self.stats_of: Dict[int, Tuple[torch.Tensor, int]] = {}
        # opt-in (env MFB_FUSE_GN_STATS=1): measured performance-neutral at the bench geometry (profiles/r01j, r01p)
        self.fuse_gn_stats = os.environ.get("MFB_FUSE_GN_STATS", "0") == "1"
        # fused BrushNet taps: (packed weight, column offset, C, zero-conv weight [C,C] f32, bias buffer, base bias, zero-conv bias)
        self.fused_taps: List[Tuple] = []
        G = cfg.norm_num_groups
        self.gn_ws = torch.zeros(ops.gn_ws_floats(B, G), device=device, dtype=f32)   # zeroed once: holds ticket counters
        ops.lib()

    # ---- memory
    def buf(self, *shape, dtype=None) -> torch.Tensor:
        t = torch.zeros(*shape, device=self.dev, dtype=dtype or self.act)
        self.keep.append(t)
        return t

    def scratch(self, role: str, *shape) -> torch.Tensor:
        key = (role,) + tuple(shape)
        if key not in self._scratch:
            self._scratch[key] = torch.zeros(*shape, device=self.dev, dtype=self.act)
        return self._scratch[key]

    def D(self, t: torch.Tensor) -> torch.Tensor:
        """A packed weight / bias -> the device (a no-op unless host_pack)."""
        return t.to(self.dev)

    def wf(self, name: str) -> torch.Tensor:
        return self.sd[name].contiguous().to(self.dev)

    # ---- BrushNet tap folded into the consuming GEMM as one more K-segment:
    #      out += s * (Wz . h_brushnet + bz)   (brushnet.py:832-834,904-906 + the tap add sites)
    def _register_fused(self, wp, koff, wz, bias_buf, base_bias, bz):
        bias_buf.copy_(base_bias + bz)          # conditioning scale 1.0 until set_tap_scale() says otherwise
        self.fused_taps.append((wp, koff, wz.shape[1], wz, bias_buf, base_bias, bz))

    def set_tap_scale(self, s: float):
        """Re-scale the fused zero-conv K-segments in place (descriptors keep pointing at the same buffers)."""
        for wp, koff, c, wz, bias_buf, base_bias, bz in self.fused_taps:
            wp[:, koff:koff + c].copy_((wz * s).to(self.act))
            bias_buf.copy_(base_bias + s * bz)

    # ---- op emitters
    def emit(self, fn: Callable[[], None], n_launch: int = 1, tag: str = "misc", flops: float = 0.0, out=None, reads=()):
        """`out` / `reads`: tensors this entry writes / reads that may belong to the OTHER network's engine — the
        two-stream scheduler of StepEngine turns them into cross-stream events."""
        pos = len(self.prog)
        self.prog.append(fn)
        self.tags.append((tag, flops))
        self.notes.append(None)
        self.exec_flops.append(flops)
        self.launches += n_launch
        if out is not None:
            self.writer_pos[out.data_ptr()] = pos
        ext = [t.data_ptr() for t in reads if t is not None]
        if ext:
            self.ext_reads[pos] = ext

    def emit_plan(self, plan: ops.ConvPlan, out=None, reads=(), gn_stats: bool = False):
        """gn_stats: the output will be normalised by a GroupNorm -> let the epilogue emit its statistics."""
        if gn_stats and self.fuse_gn_stats and out is not None:
            st = plan.enable_output_stats()
            if st is not None:
                self.stats_of[out.data_ptr()] = st
            else:
                self.stats_of.pop(out.data_ptr(), None)
        elif out is not None:
            self.stats_of.pop(out.data_ptr(), None)      # buffer rewritten by a producer without statistics
        self.keep.append(plan)
        self.flops += plan.flops
        self.emit(plan.run, getattr(plan, "launches", 1), "igemm", plan.flops, out=out, reads=reads)
        self.notes[-1] = getattr(plan, "note", None)
        self.exec_flops[-1] = getattr(plan, "exec_flops", plan.flops)

    def groupnorm(self, x1, x2, prefix: str, out, HW: int, eps: float, silu: bool):
        g, b = self.wf(prefix + ".weight"), self.wf(prefix + ".bias")
        G = self.cfg.norm_num_groups
        self.keep += [g, b]
        p1 = self.stats_of.get(x1.data_ptr())
        p2 = None if x2 is None else self.stats_of.get(x2.data_ptr())
        if p1 is None or (x2 is not None and p2 is None):
            p1 = p2 = None
        self.emit(lambda: ops.groupnorm(x1, x2, g, b, out, self.gn_ws, B=self.B, HW=HW, groups=G, eps=eps, silu=silu,
                                        part1=p1, part2=p2), 1 if HW <= 64 else 2, "groupnorm")   # <= 8x8 maps: single-launch kernel

    def layernorm(self, x, prefix: str, out):
        g, b = self.wf(prefix + ".weight"), self.wf(prefix + ".bias")
        self.keep += [g, b]
        self.emit(lambda: ops.layernorm(x, g, b, out, 1e-5), 1, "layernorm")

    # ---- timestep path (embeddings.py:27-67,226-237; resnet.py:369-376), fp32
    def build_time_path(self, resnet_prefixes: Sequence[str]):
        c0, temb = self.cfg.block_out_channels[0], self.cfg.time_embed_dim
        self.t_dev = torch.zeros(self.B, device=self.dev, dtype=f32)
        sin = self.buf(self.B, c0, dtype=f32)
        e1 = self.buf(self.B, temb, dtype=f32)
        emb = self.buf(self.B, temb, dtype=f32)
        w1, b1 = self.D(self.sd["time_embedding.linear_1.weight"].to(self.act).contiguous()), self.wf("time_embedding.linear_1.bias")
        w2, b2 = self.D(self.sd["time_embedding.linear_2.weight"].to(self.act).contiguous()), self.wf("time_embedding.linear_2.bias")
        wcat = self.D(torch.cat([self.sd[p + ".time_emb_proj.weight"] for p in resnet_prefixes], 0).to(self.act).contiguous())
        bcat = self.D(torch.cat([self.sd[p + ".time_emb_proj.bias"] for p in resnet_prefixes], 0).contiguous())
        self.rowbias = self.buf(self.B, wcat.shape[0], dtype=f32)
        self.rowbias_off = {}
        off = 0
        for p in resnet_prefixes:
            self.rowbias_off[p] = off
            off += self.sd[p + ".time_emb_proj.weight"].shape[0]
        self.keep += [w1, b1, w2, b2, wcat, bcat]
        self.emit(lambda: ops.timestep_sinusoid(self.t_dev, sin))
        self.emit(lambda: ops.linear_small(sin, w1, b1, e1, act_out=True))
        self.emit(lambda: ops.linear_small(e1, w2, b2, emb))
        self.emit(lambda: ops.linear_small(emb, wcat, bcat, self.rowbias, act_in=True))
        self.n_time_ops = len(self.prog)      # the timestep path depends only on t: a denoise loop can hoist it

    def timestep_table(self, timesteps: Sequence[float]) -> torch.Tensor:
        """Row-bias rows (all time_emb_proj outputs of this net) for a list of timesteps: [len(timesteps), sum Cout].
        Runs the timestep path B timesteps at a time, outside the step loop."""
        rows = []
        ts = [float(t) for t in timesteps]
        for i in range(0, len(ts), self.B):
            chunk = ts[i:i + self.B]
            pad = chunk + [chunk[-1]] * (self.B - len(chunk))
            self.t_dev.copy_(torch.tensor(pad, dtype=f32))
            for f in self.prog[: self.n_time_ops]:
                f()
            rows.append(self.rowbias[: len(chunk)].clone())
        return torch.cat(rows, 0)

    def run_main(self):
        """Everything after the timestep path (row biases must already be in self.rowbias)."""
        for f in self.prog[self.n_time_ops:]:
            f()

    # ---- ResnetBlock2D (resnet.py:329-405)
    def resnet(self, p: str, xa, xb, HW_hw: Tuple[int, int], cout: int, tap=None, tap_src=None):
        h, w = HW_hw
        HW = h * w
        B = self.B
        ca = xa.shape[-1]
        cb = 0 if xb is None else xb.shape[-1]
        cin = ca + cb
        eps = self.cfg.norm_eps
        n1 = self.scratch("n1", B, HW, cin)
        self.groupnorm(xa, xb, p + ".norm1", n1, HW, eps, True)
        h1 = self.scratch("h1", B, HW, cout)
        w1 = self.D(ops.pack_conv_weight(self.sd[p + ".conv1.weight"]))
        off = self.rowbias_off.get(p)          # None: a resnet without time embedding (the VAE's)
        rb = None if off is None else self.rowbias[:, off:]
        self.emit_plan(ops.ConvPlan(n1, w1, h1, B=B, H=h, W=w, Cin=cin, Cout=cout, ksize=3, bias=self.wf(p + ".conv1.bias"),
                                    rowbias=rb, rowbias_ld=0 if rb is None else self.rowbias.shape[1]), out=h1, gn_stats=True)
        n2 = self.scratch("n2", B, HW, cout)
        self.groupnorm(h1, None, p + ".norm2", n2, HW, eps, True)
        out = self.buf(B, HW, cout)
        bias = self.sd[p + ".conv2.bias"].clone()
        wmain = self.sd[p + ".conv2.weight"]
        extras_w: list = []
        extras_x: list = []
        res1 = None
        if p + ".conv_shortcut.weight" in self.sd:
            ws = self.sd[p + ".conv_shortcut.weight"][:, :, 0, 0]
            extras_w += [ws[:, :ca]] + ([ws[:, ca:]] if cb else [])
            extras_x += [xa] + ([xb] if cb else [])
            bias = bias + self.sd[p + ".conv_shortcut.bias"]
        else:
            assert cb == 0 and ca == cout
            res1 = xa
        fused = None
        if tap_src is not None:
            koff = 9 * cout + sum(e.shape[1] for e in extras_w)
            extras_w.append(tap_src[1]); extras_x.append(tap_src[0])
            fused = (koff, tap_src[1], tap_src[2])
        w2 = self.D(ops.pack_conv_weight(wmain, extras=extras_w))
        base_bias = bias.contiguous()
        bias_buf = self.D(base_bias.clone())
        if fused is not None:
            self._register_fused(w2, fused[0], fused[1], bias_buf, base_bias, fused[2])
        self.emit_plan(ops.ConvPlan(n2, w2, out, B=B, H=h, W=w, Cin=cout, Cout=cout, ksize=3, extras=extras_x,
                                    bias=bias_buf, res1=res1, res2=tap), out=out, reads=list(extras_x) + [tap], gn_stats=True)
        return out

    def _sampler_conv(self, p: str, x, out, h, w, stride, tap, tap_src):
        c = x.shape[-1]
        extras_w, extras_x = [], []
        if tap_src is not None:
            extras_w.append(tap_src[1]); extras_x.append(tap_src[0])
        wp = self.D(ops.pack_conv_weight(self.sd[p + ".conv.weight"], extras=extras_w))
        base_bias = self.sd[p + ".conv.bias"].contiguous()
        bias_buf = self.D(base_bias.clone())
        if tap_src is not None:
            self._register_fused(wp, 9 * c, tap_src[1], bias_buf, base_bias, tap_src[2])
        self.emit_plan(ops.ConvPlan(x, wp, out, B=self.B, H=h, W=w, Cin=c, Cout=c, ksize=3, stride=stride, extras=extras_x,
                                    bias=bias_buf, res2=tap), out=out, reads=list(extras_x) + [tap], gn_stats=True)

    def downsample(self, p: str, x, hw, tap=None, tap_src=None):
        h, w = hw
        c = x.shape[-1]
        out = self.buf(self.B, (h // 2) * (w // 2), c)
        self._sampler_conv(p, x, out, h, w, 2, tap, tap_src)
        return out

    def upsample(self, p: str, x, hw, tap=None, tap_src=None):
        """Upsample2D (nearest x2 + conv3x3, upsampling.py:145-186) as four sub-pixel 2x2 convolutions over the
        low-resolution tensor: no upsampled intermediate, 4/9 of the MMA work."""
        h, w = hw
        c = x.shape[-1]
        out = self.buf(self.B, 4 * h * w, c)
        extras_w, extras_x = [], []
        if tap_src is not None:
            extras_w.append(tap_src[1]); extras_x.append(tap_src[0])
        wp = self.D(ops.pack_upconv_weight(self.sd[p + ".conv.weight"], extras=extras_w))      # [4, C, 4C (+C)]
        base_bias = self.sd[p + ".conv.bias"].contiguous()
        bias_buf = self.D(base_bias.clone())
        if tap_src is not None:
            self._register_fused(wp.view(4 * c, -1), 4 * c, tap_src[1].repeat(4, 1), bias_buf, base_bias, tap_src[2])
        self.emit_plan(ops.ConvPlan(x, wp, out, B=self.B, H=h, W=w, Cin=c, Cout=c, ksize=3, up2x=True, extras=extras_x,
                                    bias=bias_buf, res2=tap), out=out, reads=list(extras_x) + [tap], gn_stats=True)
        return out

    def run(self):
        for f in self.prog:
            f()

    def run_timed(self, per_entry: bool = False, skip: int = 0):
        """Run the program once with a CUDA-event pair around every entry (on the current stream) and return
        {family: (milliseconds, algorithmic FLOPs, entries, executed FLOPs)} — the live per-kernel timing bench.py reports.
        per_entry=True returns [(family, note, ms, FLOPs)] per program entry instead (tools/profile_step.py).
        skip: leading entries to leave out (the hoisted timestep path: `n_time_ops`)."""
        evs = []
        for f in self.prog[skip:]:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            f()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if per_entry:
            return [(tag, note, a.elapsed_time(b), fl) for (a, b), (tag, fl), note in zip(evs, self.tags[skip:], self.notes[skip:])]
        out: Dict[str, List[float]] = {}
        for (a, b), (tag, fl), xf in zip(evs, self.tags[skip:], self.exec_flops[skip:]):
            r = out.setdefault(tag, [0.0, 0.0, 0, 0.0])
            r[0] += a.elapsed_time(b)
            r[1] += fl
            r[2] += 1
            r[3] += xf
        return out


def _resnet_prefixes(cfg: NetConfig) -> List[str]:
    out = []
    n = len(cfg.block_out_channels)
    for i in range(n):
        out += [f"down_blocks.{i}.resnets.{j}" for j in range(cfg.layers_per_block)]
    out += ["mid_block.resnets.0", "mid_block.resnets.1"]
    for i in range(n):
        out += [f"up_blocks.{i}.resnets.{j}" for j in range(cfg.layers_per_block + 1)]
    return out


class BrushNetEngine(_Net):
    """BrushNetModel.forward on the kernels: conv_in_condition over [latent || cond], resnet-only down/mid/up,
    28 zero-conv taps scaled by conditioning_scale (a device scalar, so the captured graph serves any scale)."""

    def __init__(self, cfg, sd, B, H, W, device, tap_bufs: Optional[List[torch.Tensor]] = None,
                 only_first_tap: bool = False, host_pack: bool = False, dup_halves: bool = False):
        """only_first_tap: fused pipeline mode — only the conv_in-site tap is materialised; the other 27 zero-convs
        are handed to the UNet engine as (feature, weight, bias) and run there as K-segments of the consuming GEMM."""
        super().__init__(cfg, sd, B, H, W, device, "brushnet", host_pack=host_pack)
        boc = cfg.block_out_channels
        n = len(boc)
        self.sample_in = torch.zeros(B, cfg.in_channels, H, W, device=device, dtype=f32)
        self.cond_in = torch.zeros(B, cfg.conditioning_channels, H, W, device=device, dtype=f32)
        # zero-conv output scales, one device scalar per tap (epilogue alpha): conditioning_scale, times logspace(-1, 0, 28) in
        # guess mode (brushnet.py:896-906).  set_scale() rewrites them; the captured graph reads them at replay.
        ntap = sum(cfg.layers_per_block + (i != len(cfg.block_out_channels) - 1) for i in range(len(cfg.block_out_channels))) + 1 + 1 + \
            sum(cfg.layers_per_block + 1 + (i != len(cfg.block_out_channels) - 1) for i in range(len(cfg.block_out_channels)))
        self.scales = torch.ones(ntap, device=device, dtype=f32)
        self._scale_base = torch.ones(ntap, dtype=f32)
        self.scale = self.scales[:1]             # the conv_in-site tap's scale (fused mode materialises only that tap)
        self.build_time_path(_resnet_prefixes(cfg))
        # conv_in_condition (brushnet.py:810-811)
        wci = self.D(self.sd["conv_in_condition.weight"].permute(2, 3, 1, 0).contiguous())
        bci = self.wf("conv_in_condition.bias")
        x = self.buf(B, H * W, boc[0])
        self.keep += [wci, bci]
        self.emit(lambda x0=x: ops.conv_in(self.sample_in, self.cond_in, wci, bci, x0), out=x)   # bind now: `x` is reassigned below
        # zero-convs (brushnet.py:831-834,851,890-893) with the conditioning scale (:904-906) as epilogue alpha.  Each one is emitted
        # RIGHT AFTER the feature it reads: a consumer on another launch stream (the UNet under StepEngine(two_streams=True)) can
        # then start as soon as its tap exists — emitted at the end of the program, the conv_in-site tap made the UNet's very first
        # kernel wait for the whole BrushNet, which serialised the two streams completely.
        self.taps: List[torch.Tensor] = []
        self.tap_hw: List[Tuple[int, int]] = []
        self.tap_sources: List[Tuple] = []

        # dup_halves (StepEngine's exact CFG de-duplication): the branch runs on b samples; every feature the UNet consumes (and the
        # conv_in-site tap) is broadcast to both CFG halves right after it is produced: dup_sources[k] / dup_tap0 are [2b, HW, C]
        self.dup_sources: List[Tuple] = []
        self.dup_tap0: Optional[torch.Tensor] = None

        def dup(t):
            d = torch.empty(2, *t.shape, device=self.dev, dtype=t.dtype)
            self.keep.append(d)
            self.emit(lambda s0=t, d0=d: d0.copy_(s0.unsqueeze(0).expand_as(d0)), out=d)
            return d.view(2 * t.shape[0], *t.shape[1:])

        def tap(src, shw, nm):
            k = len(self.tap_sources)
            wz, bz = self.sd[nm + ".weight"][:, :, 0, 0].contiguous(), self.sd[nm + ".bias"].contiguous()
            self.tap_sources.append((src, wz, bz))
            if dup_halves:
                self.dup_sources.append((src if k == 0 else dup(src), wz, bz))
            if only_first_tap and k > 0:
                return
            c = src.shape[-1]
            t = tap_bufs[k] if tap_bufs is not None else self.buf(B, shw[0] * shw[1], c)
            wz = self.D(ops.pack_conv_weight(self.sd[nm + ".weight"]))
            self.emit_plan(ops.ConvPlan(src, wz, t, B=B, H=shw[0], W=shw[1], Cin=c, Cout=c, ksize=1,
                                        bias=self.wf(nm + ".bias"), alpha=self.scales[k:k + 1]), out=t)
            self.taps.append(t)
            self.tap_hw.append(shw)
            if dup_halves and k == 0:
                self.dup_tap0 = dup(t)

        hw = (H, W)
        feats: List[Tuple[torch.Tensor, Tuple[int, int]]] = [(x, hw)]
        tap(x, hw, "brushnet_down_blocks.0")
        for i in range(n):
            for j in range(cfg.layers_per_block):
                x = self.resnet(f"down_blocks.{i}.resnets.{j}", x, None, hw, boc[i])
                feats.append((x, hw))
                tap(x, hw, f"brushnet_down_blocks.{len(feats) - 1}")
            if i != n - 1:
                x = self.downsample(f"down_blocks.{i}.downsamplers.0", x, hw)
                hw = (hw[0] // 2, hw[1] // 2)
                feats.append((x, hw))
                tap(x, hw, f"brushnet_down_blocks.{len(feats) - 1}")
        down_feats = list(feats)
        x = self.resnet("mid_block.resnets.0", x, None, hw, boc[-1])
        x = self.resnet("mid_block.resnets.1", x, None, hw, boc[-1])
        tap(x, hw, "brushnet_mid_block")
        n_up = 0
        skips = list(feats)
        for i, layers in enumerate(up_block_channels(cfg)):
            for j, (_cin, _hid, _skip, cout) in enumerate(layers):
                s, shw = skips.pop()
                assert shw == hw
                x = self.resnet(f"up_blocks.{i}.resnets.{j}", x, s, hw, cout)
                tap(x, hw, f"brushnet_up_blocks.{n_up}")
                n_up += 1
            if i != n - 1:
                x = self.upsample(f"up_blocks.{i}.upsamplers.0", x, hw)
                hw = (hw[0] * 2, hw[1] * 2)
                tap(x, hw, f"brushnet_up_blocks.{n_up}")
                n_up += 1
        self.n_down = len(down_feats)
        assert len(self.tap_sources) == ntap

    def set_scale(self, conditioning_scale: float, guess_mode: bool = False):
        """conditioning_scale for every tap; guess_mode: times torch.logspace(-1, 0, 28) (0.1 ... 1.0 from the first down tap to the
        last up tap, brushnet.py:896-902)."""
        key = (float(conditioning_scale), bool(guess_mode))
        if getattr(self, "_scale_key", None) == key:
            return
        base = torch.logspace(-1, 0, self.scales.numel(), dtype=f32) if guess_mode else self._scale_base
        self.scales.copy_(base * float(conditioning_scale))
        self._scale_key = key


class UNetEngine(_Net):
    """UNet2DConditionModel.forward (SD1.5 family) with the BrushNet taps consumed in the producing epilogues."""

    def __init__(self, cfg, sd, B, H, W, device, ctx_len: int = 77, tap_sources: Optional[List[Tuple]] = None,
                 tap0: Optional[torch.Tensor] = None, host_pack: bool = False):
        """tap_sources (fused pipeline mode): the BrushNet engine's 28 (feature, zero-conv weight, bias) triples; taps
        1..27 then run as K-segments of the consuming GEMMs and only tap 0 (conv_in site) is read as a tensor."""
        super().__init__(cfg, sd, B, H, W, device, "unet", host_pack=host_pack)
        boc = cfg.block_out_channels
        n = len(boc)
        self.ctx_len = ctx_len
        self.sample_in = torch.zeros(B, cfg.in_channels, H, W, device=device, dtype=f32)
        self.ehs_in = torch.zeros(B, ctx_len, cfg.cross_attention_dim, device=device, dtype=f32)
        self.ehs_bf = torch.zeros(B * ctx_len, cfg.cross_attention_dim, device=device, dtype=self.act)
        self.out = torch.zeros(B, cfg.out_channels, H, W, device=device, dtype=f32)
        self.ctx_prog: List[Callable[[], None]] = []
        # tap input buffers (zero == "no taps"); shapes in pop order
        dch, mch, uch = tap_channels(cfg)
        self.tap_hw: List[Tuple[int, int]] = []
        hw = (H, W)
        self.tap_hw.append(hw)
        for i in range(n):
            self.tap_hw += [hw] * cfg.layers_per_block
            if i != n - 1:
                hw = (hw[0] // 2, hw[1] // 2)
                self.tap_hw.append(hw)
        self.tap_hw.append(hw)
        for i in range(n):
            self.tap_hw += [hw] * (cfg.layers_per_block + 1)
            if i != n - 1:
                hw = (hw[0] * 2, hw[1] * 2)
                self.tap_hw.append(hw)
        chans = dch + [mch] + uch
        self.n_down = len(dch)
        if tap_sources is None:
            self.taps = [self.buf(B, h_ * w_, c) for (h_, w_), c in zip(self.tap_hw, chans)]
            tap_it = iter([(t, None) for t in self.taps])
        else:
            assert tap0 is not None and len(tap_sources) == len(chans)
            self.taps = [tap0]
            tap_it = iter([(tap0, None)] + [(None, src) for src in tap_sources[1:]])

        self.build_time_path(_resnet_prefixes(cfg))
        wci = self.D(self.sd["conv_in.weight"].permute(2, 3, 1, 0).contiguous())
        bci = self.wf("conv_in.bias")
        self.keep += [wci, bci]
        hw = (H, W)
        pre = self.buf(B, H * W, boc[0])      # first skip keeps the PRE-tap conv_in output (unet_2d_condition.py:1215-1218)
        x = self.buf(B, H * W, boc[0])
        tap0, _ = next(tap_it)
        self.emit(lambda x0=x: ops.conv_in(self.sample_in, None, wci, bci, pre, tap0, x0), reads=[tap0])  # bind now: `x` is reassigned below
        skips = [(pre, hw)]
        for i in range(n):
            for j in range(cfg.layers_per_block):
                tap, tsrc = next(tap_it)
                if cfg.down_has_attn[i]:
                    x = self.resnet(f"down_blocks.{i}.resnets.{j}", x, None, hw, boc[i])
                    x = self.transformer(f"down_blocks.{i}.attentions.{j}", x, hw, tap, tsrc)
                else:
                    x = self.resnet(f"down_blocks.{i}.resnets.{j}", x, None, hw, boc[i], tap=tap, tap_src=tsrc)
                skips.append((x, hw))
            if i != n - 1:
                tap, tsrc = next(tap_it)
                x = self.downsample(f"down_blocks.{i}.downsamplers.0", x, hw, tap=tap, tap_src=tsrc)
                hw = (hw[0] // 2, hw[1] // 2)
                skips.append((x, hw))
        # mid (UNetMidBlock2DCrossAttn unet_2d_blocks.py:850-899) + mid tap (unet_2d_condition.py:1288-1289)
        x = self.resnet("mid_block.resnets.0", x, None, hw, boc[-1])
        x = self.transformer("mid_block.attentions.0", x, hw, None, None)
        tap, tsrc = next(tap_it)
        x = self.resnet("mid_block.resnets.1", x, None, hw, boc[-1], tap=tap, tap_src=tsrc)
        for i, layers in enumerate(up_block_channels(cfg)):
            for j, (_cin, _hid, _skip, cout) in enumerate(layers):
                s, shw = skips.pop()
                assert shw == hw
                tap, tsrc = next(tap_it)
                if cfg.up_has_attn[i]:
                    x = self.resnet(f"up_blocks.{i}.resnets.{j}", x, s, hw, cout)
                    x = self.transformer(f"up_blocks.{i}.attentions.{j}", x, hw, tap, tsrc)
                else:
                    x = self.resnet(f"up_blocks.{i}.resnets.{j}", x, s, hw, cout, tap=tap, tap_src=tsrc)
            if i != n - 1:
                tap, tsrc = next(tap_it)
                x = self.upsample(f"up_blocks.{i}.upsamplers.0", x, hw, tap=tap, tap_src=tsrc)
                hw = (hw[0] * 2, hw[1] * 2)
        # conv_norm_out -> SiLU -> conv_out (unet_2d_condition.py:1336-1339)
        nout = self.scratch("n1", B, H * W, boc[0])
        self.groupnorm(x, None, "conv_norm_out", nout, H * W, cfg.norm_eps, True)
        wco = self.D(self.sd["conv_out.weight"].permute(0, 2, 3, 1).contiguous())
        bco = self.wf("conv_out.bias")
        self.keep += [wco, bco]
        self.emit(lambda: ops.conv_out(nout, wco, bco, self.out, B=B, H=H, W=W))  # nout/wco/bco are not rebound

    # Transformer2DModel + BasicTransformerBlock (transformer_2d.py:334-430, attention.py:291-412)
    def transformer(self, p: str, x, hw, tap, tap_src=None):
        B, cfg = self.B, self.cfg
        h, w = hw
        T = h * w
        M = B * T
        C = x.shape[-1]
        heads = cfg.heads
        d = C // heads
        t = p + ".transformer_blocks.0"
        g = self.scratch("tg", B, T, C)
        self.groupnorm(x, None, p + ".norm", g, T, 1e-6, False)
        h0 = self.scratch("th0", M, C)
        self.emit_plan(ops.linear_plan(g.view(M, C), self.D(ops.pack_conv_weight(self.sd[p + ".proj_in.weight"])), h0,
                                       bias=self.wf(p + ".proj_in.bias")))
        # --- self attention
        nrm = self.scratch("tn", M, C)
        self.layernorm(h0, t + ".norm1", nrm)
        qkv = self.scratch("tqkv", M, 3 * C)
        wqkv = self.D(torch.cat([self.sd[t + ".attn1.to_q.weight"], self.sd[t + ".attn1.to_k.weight"],
                                 self.sd[t + ".attn1.to_v.weight"]], 0).to(self.act).contiguous())
        self.emit_plan(ops.linear_plan(nrm, wqkv, qkv))
        att = self.scratch("tatt", M, C)
        kview, vview = qkv.view(-1)[C:], qkv.view(-1)[2 * C:]      # q | k | v column blocks of the fused projection
        self.emit(lambda: ops.attention(qkv, kview, vview, att, B=B, heads=heads, head_dim=d, Tq=T, Tk=T, ldq=3 * C,
                                        ldk=3 * C, ldv=3 * C, ldo=C), 1, "attention", 4.0 * B * T * T * C)
        self.flops += 4.0 * B * T * T * C
        h1 = self.scratch("th1", M, C)
        self.emit_plan(ops.linear_plan(att, self.D(self.sd[t + ".attn1.to_out.0.weight"].to(self.act).contiguous()), h1,
                                       bias=self.wf(t + ".attn1.to_out.0.bias"), res1=h0))
        # --- cross attention (K/V of the context are prepared once per prompt: ctx_prog)
        Lc = self.ctx_len
        k2 = self.buf(B * Lc, C)
        v2 = self.buf(B * Lc, C)
        pk = ops.linear_plan(self.ehs_bf, self.D(self.sd[t + ".attn2.to_k.weight"].to(self.act).contiguous()), k2)
        pv = ops.linear_plan(self.ehs_bf, self.D(self.sd[t + ".attn2.to_v.weight"].to(self.act).contiguous()), v2)
        self.keep += [pk, pv]
        self.ctx_prog += [pk.run, pv.run]
        self.layernorm(h1, t + ".norm2", nrm)
        q2 = self.scratch("tq2", M, C)
        self.emit_plan(ops.linear_plan(nrm, self.D(self.sd[t + ".attn2.to_q.weight"].to(self.act).contiguous()), q2))
        self.emit(lambda: ops.attention(q2, k2, v2, att, B=B, heads=heads, head_dim=d, Tq=T, Tk=Lc, ldq=C, ldk=C,
                                        ldv=C, ldo=C), 1, "attention", 4.0 * B * T * Lc * C)
        self.flops += 4.0 * B * T * Lc * C
        h2 = self.scratch("th2", M, C)
        self.emit_plan(ops.linear_plan(att, self.D(self.sd[t + ".attn2.to_out.0.weight"].to(self.act).contiguous()), h2,
                                       bias=self.wf(t + ".attn2.to_out.0.bias"), res1=h1))
        # --- GEGLU feed-forward
        self.layernorm(h2, t + ".norm3", nrm)
        wg, bg = (self.D(a) for a in ops.pack_geglu(self.sd[t + ".ff.net.0.proj.weight"], self.sd[t + ".ff.net.0.proj.bias"]))
        gg = self.scratch("tgg", M, 4 * C)
        self.emit_plan(ops.linear_plan(nrm, wg, gg, bias=bg, geglu=True))
        h3 = self.scratch("th3", M, C)
        self.emit_plan(ops.linear_plan(gg, self.D(self.sd[t + ".ff.net.2.weight"].to(self.act).contiguous()), h3,
                                       bias=self.wf(t + ".ff.net.2.bias"), res1=h2))
        # --- proj_out + transformer residual (+ BrushNet tap, added after the attention: unet_2d_blocks.py:1374-1389)
        out = self.buf(B, T, C)
        extras_w, extras_x = [], []
        if tap_src is not None:
            extras_w.append(tap_src[1]); extras_x.append(tap_src[0].view(M, C))
        wpo = self.D(ops.pack_conv_weight(self.sd[p + ".proj_out.weight"], extras=extras_w))
        base_bias = self.sd[p + ".proj_out.bias"].contiguous()
        bias_buf = self.D(base_bias.clone())
        if tap_src is not None:
            self._register_fused(wpo, C, tap_src[1], bias_buf, base_bias, tap_src[2])
        # a 1x1 conv with the real (B, h, w) geometry rather than a flat token GEMM: the epilogue's fused GroupNorm
        # statistics are per image
        self.emit_plan(ops.ConvPlan(h3, wpo, out, B=B, H=h, W=w, Cin=h3.shape[-1], Cout=C, ksize=1,
                                    extras=[e.view(B, T, -1) for e in extras_x], bias=bias_buf, res1=x, res2=tap),
                       out=out, reads=list(extras_x) + [tap], gn_stats=True)
        return out

    def set_context(self, ehs: torch.Tensor):
        """encoder_hidden_states [B, 77, ctx] -> cached cross-attention K / V^T of all 16 layers."""
        self.ehs_in.copy_(ehs.to(device=self.dev, dtype=f32))
        ops.f32_to_bf16(self.ehs_in, self.ehs_bf)
        for f in self.ctx_prog:
            f()
