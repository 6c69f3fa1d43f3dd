"""Seeded synthetic weights and SynMirror-shaped inputs.

There is no network for checkpoints, so benches and parity tests use random-init
weights of the SD1.5 architecture (BASELINE.json `configs`).  Every tensor is drawn
from its own generator seeded by (seed, crc32(name)), so the same `state_dict` is
reproduced on any machine and can be loaded (strict) into the reference modules
(`UNet2DConditionModel.load_state_dict`, `BrushNetModel.load_state_dict`) to make
golden vectors.  Distributions follow torch's default layer init (uniform ±1/sqrt(fan_in));
the 28 BrushNet zero-convs get normal(std=0.02) instead of zeros so that the taps
are exercised (reference zero-init is S/models/brushnet.py:928-931).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import torch

from .config import NetConfig, param_shapes


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFF)
    return g


def make_state_dict(cfg: NetConfig, net: str, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg, net):
        g = _gen(seed, net + "/" + name)
        is_norm = ".norm" in name or name.startswith("conv_norm_out")
        if name.startswith("brushnet_") and name.endswith(".weight"):
            t = torch.randn(shape, generator=g) * 0.02
        elif name.startswith("brushnet_") and name.endswith(".bias"):
            t = torch.randn(shape, generator=g) * 0.01
        elif is_norm and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            if name.endswith(".weight"):
                fan_in = math.prod(shape[1:])
            else:  # bias: fan_in of the matching weight is not known here; use a small fixed bound
                fan_in = 1024
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        sd[name] = t.to(dtype)
    return sd


def make_inputs(cfg: NetConfig, images: int, seed: int = 1234, height: int | None = None, width: int | None = None,
                cfg_duplicate: bool = True):
    """Latent-space inputs of one denoise call, shaped like the reference pipeline builds them
    (S/pipelines/brushnet/pipeline_brushnet.py:1176-1202): initial latents [b,4,h,w],
    conditioning_latents [2b,6,h,w] = masked-image latent x0.18215 || mask {0,1} || depth in [-1,1],
    prompt_embeds [2b,77,ctx] with the unconditional half first (:1102-1103)."""
    h = height or cfg.sample_size
    w = width or cfg.sample_size
    g = torch.Generator(device="cpu").manual_seed(seed)
    lat = torch.randn(images, cfg.in_channels, h, w, generator=g)
    masked = torch.randn(images, 4, h, w, generator=g) * 0.18215 * 4.0
    # filled axis-aligned rectangle covering ~20-40 % of the frame is the mirror region (mask = 0 there)
    mask = torch.ones(images, 1, h, w)
    for i in range(images):
        fh = int(h * (0.45 + 0.15 * torch.rand((), generator=g).item()))
        fw = int(w * (0.45 + 0.15 * torch.rand((), generator=g).item()))
        y0 = int((h - fh) * torch.rand((), generator=g).item())
        x0 = int((w - fw) * torch.rand((), generator=g).item())
        mask[i, :, y0:y0 + fh, x0:x0 + fw] = 0.0
    yy = torch.linspace(-1, 1, h).view(1, 1, h, 1)
    xx = torch.linspace(-1, 1, w).view(1, 1, 1, w)
    a = torch.rand(images, 1, 1, 1, generator=g) - 0.5
    b = torch.rand(images, 1, 1, 1, generator=g) - 0.5
    depth = (a * yy + b * xx + 0.3 * torch.exp(-4 * (yy * yy + xx * xx))).clamp(-1, 1)
    cond = torch.cat([masked, mask, depth], 1)[:, : cfg.conditioning_channels]
    neg = torch.randn(images, 77, cfg.cross_attention_dim, generator=g)
    pos = torch.randn(images, 77, cfg.cross_attention_dim, generator=g)
    if cfg_duplicate:
        cond = torch.cat([cond, cond], 0)
        ehs = torch.cat([neg, pos], 0)
    else:
        ehs = pos
    return {"latents": lat, "conditioning_latents": cond, "prompt_embeds": ehs}
