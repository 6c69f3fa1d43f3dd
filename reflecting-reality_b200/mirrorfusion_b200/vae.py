"""AutoencoderKL.decode on the libmfb200 kernels (once the 50-step loop takes ~1.2 s per 8
images, the torch VAE decode at 512x512 is the visible tail of `images/s`).

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


def make_vae_state_dict(cfg: VaeConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic decoder weights (same rule as synth.make_state_dict), loadable into the reference AutoencoderKL."""
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in vae_decoder_param_shapes(cfg):
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


class VaeDecoderEngine(_Net):
    """decode(z) == AutoencoderKL.decode(z).sample for z [B, 4, h, w] (the caller divides by scaling_factor, as
    pipeline_brushnet.py:1337 does): returns [B, 3, 8h, 8w] fp32 (for the 4-level SD VAE)."""

    def __init__(self, cfg: VaeConfig, state_dict: Dict[str, torch.Tensor], B: int, h: int, w: int, device="cuda"):
        dec = {k[len("decoder."):]: v for k, v in state_dict.items() if k.startswith("decoder.")}
        super().__init__(cfg, dec, B, h, w, device, "vae_decoder")
        self.rowbias, self.rowbias_off = None, {}                      # no time embedding in the VAE resnets
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
        self.flops += 0.0
        out = self.buf(B, T, C)
        self.emit_plan(ops.linear_plan(att, self.sd[p + ".to_out.0.weight"].to(act).contiguous(), out.view(M, C),
                                       bias=self.wf(p + ".to_out.0.bias"), res1=x.view(M, C)), out=out)
        return out

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        self.z_in.copy_(z.to(device=self.dev, dtype=f32))
        self.run()
        return self.out
