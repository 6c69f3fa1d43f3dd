"""AutoencoderKL.decode and .encode on the libmfb200 kernels (once the 50-step loop takes ~1.2 s
per 8 images, the torch VAE at 512x512 — one decode and one or two encodes per image — is the visible tail of `images/s`).

Reference: AutoencoderKL.decode (S/models/autoencoders/autoencoder_kl.py:280-309: post_quant_conv -> Decoder),
Decoder.forward (S/models/autoencoders/vae.py:284-349), UNetMidBlock2D (S/models/unets/unet_2d_blocks.py:595-770) with
its single-head Attention (group_norm, q/k/v WITH bias, residual_connection, S/models/attention_processor.py:1204-1286),
UpDecoderBlock2D (unet_2d_blocks.py:2769-2842), ResnetBlock2D without a time embedding (S/models/resnet.py:329-405).

Everything is reused from the denoise path: the tcgen05 implicit-GEMM conv (3x3, shortcut 1x1 as K-segments, sub-pixel
upsample convs), the GroupNorm+SiLU kernel (4 channels per group at the 128-channel level), conv_in / conv_out boundary
kernels.  Two things are specific:
  * post_quant_conv (1x1, 4 -> 4) is folded EXACTLY into conv_in: conv_in(W_pq z + b_pq) with zero padding equals one 3x3
    conv over [z, 1] (a constant-ones fifth channel carries b_pq, and is zero-padded at the border exactly like W_pq z + b_pq
    is in the reference);
  * the encoder's Downsample2D(padding=0) pads one zero row / column at the bottom / right only
    (S/models/downsampling.py:141-143): the stride-2 conv plan's `pad0` mode; its conv_out (512 -> 8) and quant_conv (1x1,
    8 -> 8; S/models/autoencoders/autoencoder_kl.py:262) are folded into ONE 3x3 conv (exact: the 1x1 follows the 3x3) that
    the fp32-output boundary kernel evaluates as the mean and the logvar halves, so the latent moments never round to bf16;
  * the mid-block attention has ONE head of dim 512 (256 KB of Q and K per 128-row tile: no flash tiling fits); it runs
    per image as two tensor-core GEMMs (S = Q K^T * scale, O = P V with the transposed V as the "weight") around a
    row-softmax kernel.  In fp32 parity mode the CUDA-core attention kernel handles it directly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

from . import ops
from .engine import _Net
from .synth import _gen

f32 = torch.float32


@dataclass(frozen=True)
class VaeConfig:
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)      # encoder order, as in AutoencoderKL's config
    layers_per_block: int = 2
    latent_channels: int = 4
    out_channels: int = 3
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215


SD_VAE = VaeConfig()                                                  # the SD1.5 VAE MirrorFusion ships with
TINY_VAE = VaeConfig(block_out_channels=(128, 256), layers_per_block=1)


def vae_decoder_param_shapes(cfg: VaeConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Names/shapes of AutoencoderKL.state_dict() restricted to post_quant_conv.* and decoder.* (reference layout)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    lc = cfg.latent_channels
    out += [("post_quant_conv.weight", (lc, lc, 1, 1)), ("post_quant_conv.bias", (lc,))]
    rev = list(reversed(cfg.block_out_channels))
    cm = rev[0]
    out += [("decoder.conv_in.weight", (cm, lc, 3, 3)), ("decoder.conv_in.bias", (cm,))]

    def resnet(p, cin, cout):
        r = [(f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)), (f"{p}.conv1.weight", (cout, cin, 3, 3)),
             (f"{p}.conv1.bias", (cout,)), (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)),
             (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,))]
        if cin != cout:
            r += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]
        return r

    a = "decoder.mid_block.attentions.0"
    out += resnet("decoder.mid_block.resnets.0", cm, cm)
    out += [(f"{a}.group_norm.weight", (cm,)), (f"{a}.group_norm.bias", (cm,))]
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        out += [(f"{a}.{nm}.weight", (cm, cm)), (f"{a}.{nm}.bias", (cm,))]
    out += resnet("decoder.mid_block.resnets.1", cm, cm)
    prev = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            out += resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i != len(rev) - 1:
            out += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (c, c, 3, 3)),
                    (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (c,))]
        prev = c
    c0 = cfg.block_out_channels[0]
    out += [("decoder.conv_norm_out.weight", (c0,)), ("decoder.conv_norm_out.bias", (c0,)),
            ("decoder.conv_out.weight", (cfg.out_channels, c0, 3, 3)), ("decoder.conv_out.bias", (cfg.out_channels,))]
    return out


def vae_encoder_param_shapes(cfg: VaeConfig, in_channels: int = 3) -> List[Tuple[str, Tuple[int, ...]]]:
    """Names/shapes of AutoencoderKL.state_dict() restricted to encoder.* and quant_conv.* (reference layout)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    boc = cfg.block_out_channels
    lc2 = 2 * cfg.latent_channels

    def resnet(p, cin, cout):
        r = [(f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)), (f"{p}.conv1.weight", (cout, cin, 3, 3)),
             (f"{p}.conv1.bias", (cout,)), (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)),
             (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,))]
        if cin != cout:
            r += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]
        return r

    out += [("encoder.conv_in.weight", (boc[0], in_channels, 3, 3)), ("encoder.conv_in.bias", (boc[0],))]
    prev = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            out += resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i != len(boc) - 1:
            out += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (c, c, 3, 3)),
                    (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (c,))]
        prev = c
    cm = boc[-1]
    a = "encoder.mid_block.attentions.0"
    out += resnet("encoder.mid_block.resnets.0", cm, cm)
    out += [(f"{a}.group_norm.weight", (cm,)), (f"{a}.group_norm.bias", (cm,))]
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        out += [(f"{a}.{nm}.weight", (cm, cm)), (f"{a}.{nm}.bias", (cm,))]
    out += resnet("encoder.mid_block.resnets.1", cm, cm)
    out += [("encoder.conv_norm_out.weight", (cm,)), ("encoder.conv_norm_out.bias", (cm,)),
            ("encoder.conv_out.weight", (lc2, cm, 3, 3)), ("encoder.conv_out.bias", (lc2,)),
            ("quant_conv.weight", (lc2, lc2, 1, 1)), ("quant_conv.bias", (lc2,))]
    return out


def make_vae_state_dict(cfg: VaeConfig, seed: int = 0, part: str = "decoder") -> Dict[str, torch.Tensor]:
    """Seeded synthetic VAE weights (same rule as synth.make_state_dict), loadable into the reference AutoencoderKL.
    part: "decoder" (post_quant_conv + decoder), "encoder" (encoder + quant_conv) or "both"."""
    sd: Dict[str, torch.Tensor] = {}
    shapes = ([] if part == "encoder" else vae_decoder_param_shapes(cfg)) + ([] if part == "decoder" else vae_encoder_param_shapes(cfg))
    for name, shape in shapes:
        g = _gen(seed, "vae/" + name)
        is_norm = "norm" in name.split(".")[-2]
        if is_norm and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm:
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:]) if name.endswith(".weight") else 1024
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[name] = t
    return sd


class _VaeNet(_Net):
    """Shared by the decoder and the encoder: resnets without a time embedding and the single-head mid-block attention."""

    def __init__(self, cfg, sd, B, h, w, device, name):
        super().__init__(cfg, sd, B, h, w, device, name)
        self.rowbias, self.rowbias_off = None, {}                      # no time embedding in the VAE resnets

    def attention_block(self, p: str, x, hw):
        B, cfg = self.B, self.cfg
        T = hw[0] * hw[1]
        M = B * T
        C = x.shape[-1]
        act = self.act
        g = self.scratch("ag", B, T, C)
        self.groupnorm(x, None, p + ".group_norm", g, T, cfg.norm_eps, False)
        q, k, v = self.buf(M, C), self.buf(M, C), self.buf(M, C)
        for nm, dst in (("to_q", q), ("to_k", k), ("to_v", v)):
            self.emit_plan(ops.linear_plan(g.view(M, C), self.sd[f"{p}.{nm}.weight"].to(act).contiguous(), dst,
                                           bias=self.wf(f"{p}.{nm}.bias")))
        att = self.buf(M, C)
        if act == f32 or C <= 160:
            self.emit(lambda: ops.attention(q, k, v, att, B=B, heads=1, head_dim=C, Tq=T, Tk=T), 1, "attention", 4.0 * B * T * T * C)
            self.flops += 4.0 * B * T * T * C
        else:
            vt = self.buf(B, C, T)
            self.emit(lambda: ops.transpose_tokens(v, vt, ld=C, col0=0, Cc=C, B=B, T=T, ldt=T))
            s = self.scratch("as", T, T)
            scale = torch.full((1,), float(C) ** -0.5, device=self.dev, dtype=f32)
            self.keep.append(scale)
            for b in range(B):
                rows = slice(b * T, (b + 1) * T)
                self.emit_plan(ops.linear_plan(q[rows], k[rows], s, alpha=scale))          # S = Q K^T * C^-0.5
                self.emit(lambda: ops.softmax_rows(s, s), 1, "misc")
                self.emit_plan(ops.linear_plan(s, vt[b], att[rows]))                       # O = P V
        out = self.buf(B, T, C)
        self.emit_plan(ops.linear_plan(att, self.sd[p + ".to_out.0.weight"].to(act).contiguous(), out.view(M, C),
                                       bias=self.wf(p + ".to_out.0.bias"), res1=x.view(M, C)), out=out)
        return out


class VaeDecoderEngine(_VaeNet):
    """decode(z) == AutoencoderKL.decode(z).sample for z [B, 4, h, w] (the caller divides by scaling_factor, as
    pipeline_brushnet.py:1337 does): returns [B, 3, 8h, 8w] fp32 (for the 4-level SD VAE)."""

    def __init__(self, cfg: VaeConfig, state_dict: Dict[str, torch.Tensor], B: int, h: int, w: int, device="cuda"):
        dec = {k[len("decoder."):]: v for k, v in state_dict.items() if k.startswith("decoder.")}
        super().__init__(cfg, dec, B, h, w, device, "vae_decoder")
        dev = self.dev
        lc = cfg.latent_channels
        rev = list(reversed(cfg.block_out_channels))
        cm = rev[0]
        self.z_in = torch.zeros(B, lc, h, w, device=dev, dtype=f32)
        ones = torch.ones(B, 1, h, w, device=dev, dtype=f32)
        # post_quant_conv folded into conv_in over [z, 1] (module docstring)
        wpq = state_dict["post_quant_conv.weight"].to(dev, f32)[:, :, 0, 0]            # [j, i]
        bpq = state_dict["post_quant_conv.bias"].to(dev, f32)
        wc = self.sd["conv_in.weight"]                                                  # [o, j, kh, kw]
        wz = torch.einsum("ojhw,ji->oihw", wc, wpq)
        w1 = torch.einsum("ojhw,j->ohw", wc, bpq)[:, None]
        wci = torch.cat([wz, w1], 1).permute(2, 3, 1, 0).contiguous()                  # [3, 3, lc + 1, cm]
        bci = self.wf("conv_in.bias")
        x = self.buf(B, h * w, cm)
        self.keep += [ones, wci, bci]
        self.emit(lambda x0=x: ops.conv_in(self.z_in, ones, wci, bci, x0), out=x)
        hw = (h, w)
        # mid block: resnet, single-head attention, resnet
        x = self.resnet("mid_block.resnets.0", x, None, hw, cm)
        x = self.attention_block("mid_block.attentions.0", x, hw)
        x = self.resnet("mid_block.resnets.1", x, None, hw, cm)
        # up blocks
        for i, c in enumerate(rev):
            for j in range(cfg.layers_per_block + 1):
                x = self.resnet(f"up_blocks.{i}.resnets.{j}", x, None, hw, c)
            if i != len(rev) - 1:
                x = self.upsample(f"up_blocks.{i}.upsamplers.0", x, hw)
                hw = (hw[0] * 2, hw[1] * 2)
        self.out_hw = hw
        c0 = cfg.block_out_channels[0]
        nout = self.scratch("n1", B, hw[0] * hw[1], c0)
        self.groupnorm(x, None, "conv_norm_out", nout, hw[0] * hw[1], cfg.norm_eps, True)
        wco = self.sd["conv_out.weight"].permute(0, 2, 3, 1).contiguous()
        bco = self.wf("conv_out.bias")
        self.out = torch.zeros(B, cfg.out_channels, hw[0], hw[1], device=dev, dtype=f32)
        self.keep += [wco, bco]
        self.emit(lambda: ops.conv_out(nout, wco, bco, self.out, B=B, H=hw[0], W=hw[1]))

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        self.z_in.copy_(z.to(device=self.dev, dtype=f32))
        self.run()
        return self.out


class VaeEncoderEngine(_VaeNet):
    """encode(x) == AutoencoderKL.encode(x).latent_dist for x [B, 3, H, W] in [-1, 1]: fills `mean` / `logvar`
    ([B, 4, H/8, W/8] fp32 for the 4-level SD VAE) and returns scale * sample (`noise=None`: scale * mode), i.e. with
    scale = scaling_factor the `conditioning_latents` part of pipeline_brushnet.py:1188-1192."""

    def __init__(self, cfg: VaeConfig, state_dict: Dict[str, torch.Tensor], B: int, H: int, W: int, device="cuda"):
        enc = {k[len("encoder."):]: v for k, v in state_dict.items() if k.startswith("encoder.")}
        super().__init__(cfg, enc, B, H, W, device, "vae_encoder")
        dev = self.dev
        boc = cfg.block_out_channels
        lc = cfg.latent_channels
        self.x_in = torch.zeros(B, self.sd["conv_in.weight"].shape[1], H, W, device=dev, dtype=f32)
        wci = self.sd["conv_in.weight"].permute(2, 3, 1, 0).contiguous()
        bci = self.wf("conv_in.bias")
        x = self.buf(B, H * W, boc[0])
        self.keep += [wci, bci]
        self.emit(lambda x0=x: ops.conv_in(self.x_in, None, wci, bci, x0), out=x)
        hw = (H, W)
        for i, c in enumerate(boc):
            for j in range(cfg.layers_per_block):
                x = self.resnet(f"down_blocks.{i}.resnets.{j}", x, None, hw, c)
            if i != len(boc) - 1:
                p = f"down_blocks.{i}.downsamplers.0"
                out = self.buf(B, (hw[0] // 2) * (hw[1] // 2), c)
                self.emit_plan(ops.ConvPlan(x, ops.pack_conv_weight(self.sd[p + ".conv.weight"]), out, B=B, H=hw[0], W=hw[1], Cin=c,
                                            Cout=c, ksize=3, stride=2, pad0=True, bias=self.wf(p + ".conv.bias")), out=out)
                x, hw = out, (hw[0] // 2, hw[1] // 2)
        cm = boc[-1]
        x = self.resnet("mid_block.resnets.0", x, None, hw, cm)
        x = self.attention_block("mid_block.attentions.0", x, hw)
        x = self.resnet("mid_block.resnets.1", x, None, hw, cm)
        nout = self.scratch("n1", B, hw[0] * hw[1], cm)
        self.groupnorm(x, None, "conv_norm_out", nout, hw[0] * hw[1], cfg.norm_eps, True)
        # conv_out (3x3, cm -> 2 lc) followed by quant_conv (1x1): one 3x3 conv with composed weights, evaluated by the
        # fp32-output boundary kernel as two 4-channel halves (mean | logvar)
        wq = state_dict["quant_conv.weight"].to(dev, f32)[:, :, 0, 0]                   # [o, j]
        bq = state_dict["quant_conv.bias"].to(dev, f32)
        wc = torch.einsum("oj,jchw->ochw", wq, self.sd["conv_out.weight"])              # [2 lc, cm, 3, 3]
        bc = wq @ self.sd["conv_out.bias"] + bq
        wc = wc.permute(0, 2, 3, 1).contiguous()                                        # [Cout, 3, 3, Cin]
        self.mean = torch.zeros(B, lc, hw[0], hw[1], device=dev, dtype=f32)
        self.logvar = torch.zeros_like(self.mean)
        self.latent = torch.zeros_like(self.mean)
        for half, dst in ((0, self.mean), (1, self.logvar)):
            wh, bh = wc[half * lc:(half + 1) * lc].contiguous(), bc[half * lc:(half + 1) * lc].contiguous()
            self.keep += [wh, bh]
            self.emit(lambda w_=wh, b_=bh, d_=dst: ops.conv_out(nout, w_, b_, d_, B=B, H=hw[0], W=hw[1]))
        self.latent_hw = hw

    def encode(self, x: torch.Tensor, noise: torch.Tensor = None, scale: float = 1.0) -> torch.Tensor:
        self.x_in.copy_(x.to(device=self.dev, dtype=f32))
        self.run()
        nz = None if noise is None else noise.to(device=self.dev, dtype=f32).contiguous()
        ops.latent_sample(self.mean, self.logvar, nz, scale, self.latent)
        return self.latent
