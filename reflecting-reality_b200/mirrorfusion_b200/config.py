"""Static description of the MirrorFusion denoising networks (SD1.5 UNet + BrushNet branch).

The reference builds these from `@register_to_config` objects
(S/models/unets/unet_2d_condition.py:167-223, S/models/brushnet.py:138-214).  Only the
SD1.5-shaped family is on the hot path, so the description here is a plain dataclass
plus helpers that enumerate every parameter tensor (name, shape) exactly as
`state_dict()` of the reference modules names them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class NetConfig:
    in_channels: int = 4
    out_channels: int = 4
    conditioning_channels: int = 6          # 4 masked-image latent + 1 mask + 1 depth
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    # SD1.5: three cross-attention levels then a plain one (unet_2d_condition.py:170-176)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)
    heads: int = 8                           # `attention_head_dim=8` means 8 heads (unet_2d_condition.py:231-237)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    sample_size: int = 64

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


SD15 = NetConfig()

# A small member of the same family (channel counts stay multiples of 64 so the
# tcgen05 path applies); used by fast parity tests.
TINY = NetConfig(block_out_channels=(64, 128, 128, 128), heads=2, cross_attention_dim=64, sample_size=16,
                  norm_num_groups=8)
# An even smaller one for CPU-only oracle pins against the reference (any channel count).
MICRO = NetConfig(block_out_channels=(32, 64, 64, 64), heads=2, cross_attention_dim=32, sample_size=8)


def _resnet(p: str, cin: int, cout: int, temb: int) -> List[Tuple[str, Tuple[int, ...]]]:
    out = [
        (f"{p}.norm1.weight", (cin,)), (f"{p}.norm1.bias", (cin,)),
        (f"{p}.conv1.weight", (cout, cin, 3, 3)), (f"{p}.conv1.bias", (cout,)),
        (f"{p}.time_emb_proj.weight", (cout, temb)), (f"{p}.time_emb_proj.bias", (cout,)),
        (f"{p}.norm2.weight", (cout,)), (f"{p}.norm2.bias", (cout,)),
        (f"{p}.conv2.weight", (cout, cout, 3, 3)), (f"{p}.conv2.bias", (cout,)),
    ]
    if cin != cout:
        out += [(f"{p}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{p}.conv_shortcut.bias", (cout,))]
    return out


def _transformer(p: str, c: int, ctx: int) -> List[Tuple[str, Tuple[int, ...]]]:
    t = f"{p}.transformer_blocks.0"
    return [
        (f"{p}.norm.weight", (c,)), (f"{p}.norm.bias", (c,)),
        (f"{p}.proj_in.weight", (c, c, 1, 1)), (f"{p}.proj_in.bias", (c,)),
        (f"{t}.norm1.weight", (c,)), (f"{t}.norm1.bias", (c,)),
        (f"{t}.attn1.to_q.weight", (c, c)), (f"{t}.attn1.to_k.weight", (c, c)), (f"{t}.attn1.to_v.weight", (c, c)),
        (f"{t}.attn1.to_out.0.weight", (c, c)), (f"{t}.attn1.to_out.0.bias", (c,)),
        (f"{t}.norm2.weight", (c,)), (f"{t}.norm2.bias", (c,)),
        (f"{t}.attn2.to_q.weight", (c, c)), (f"{t}.attn2.to_k.weight", (c, ctx)), (f"{t}.attn2.to_v.weight", (c, ctx)),
        (f"{t}.attn2.to_out.0.weight", (c, c)), (f"{t}.attn2.to_out.0.bias", (c,)),
        (f"{t}.norm3.weight", (c,)), (f"{t}.norm3.bias", (c,)),
        (f"{t}.ff.net.0.proj.weight", (8 * c, c)), (f"{t}.ff.net.0.proj.bias", (8 * c,)),
        (f"{t}.ff.net.2.weight", (c, 4 * c)), (f"{t}.ff.net.2.bias", (c,)),
        (f"{p}.proj_out.weight", (c, c, 1, 1)), (f"{p}.proj_out.bias", (c,)),
    ]


def up_block_channels(cfg: NetConfig):
    """Per up block: list of (resnet_in_channels = hidden + skip, hidden_c, skip_c, out_c).

    Mirrors the channel bookkeeping of `get_up_block` callers
    (unet_2d_condition.py:520-566; UpBlock2D.__init__ unet_2d_blocks.py:2660-2680).
    """
    boc = cfg.block_out_channels
    rev = list(reversed(boc))
    n = len(boc)
    blocks = []
    output_channel = rev[0]
    for i in range(n):
        prev_output_channel = output_channel
        output_channel = rev[i]
        input_channel = rev[min(i + 1, n - 1)]
        layers = []
        nl = cfg.layers_per_block + 1
        for j in range(nl):
            res_skip = input_channel if j == nl - 1 else output_channel
            hid = prev_output_channel if j == 0 else output_channel
            layers.append((hid + res_skip, hid, res_skip, output_channel))
        blocks.append(layers)
    return blocks


def param_shapes(cfg: NetConfig, net: str) -> List[Tuple[str, Tuple[int, ...]]]:
    """Every parameter of `UNet2DConditionModel` (net="unet") or `BrushNetModel`
    (net="brushnet") in `state_dict()` naming, with its shape."""
    assert net in ("unet", "brushnet")
    boc = cfg.block_out_channels
    temb = cfg.time_embed_dim
    ctx = cfg.cross_attention_dim
    bn = net == "brushnet"
    out: List[Tuple[str, Tuple[int, ...]]] = []
    if bn:
        out += [("conv_in_condition.weight", (boc[0], cfg.in_channels + cfg.conditioning_channels, 3, 3)),
                ("conv_in_condition.bias", (boc[0],))]
    else:
        out += [("conv_in.weight", (boc[0], cfg.in_channels, 3, 3)), ("conv_in.bias", (boc[0],))]
    out += [("time_embedding.linear_1.weight", (temb, boc[0])), ("time_embedding.linear_1.bias", (temb,)),
            ("time_embedding.linear_2.weight", (temb, temb)), ("time_embedding.linear_2.bias", (temb,))]
    # down
    ch = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            out += _resnet(f"down_blocks.{i}.resnets.{j}", ch, c, temb)
            ch = c
            if cfg.down_has_attn[i] and not bn:
                out += _transformer(f"down_blocks.{i}.attentions.{j}", c, ctx)
        if i != len(boc) - 1:
            out += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (c, c, 3, 3)),
                    (f"down_blocks.{i}.downsamplers.0.conv.bias", (c,))]
    # mid
    c = boc[-1]
    out += _resnet("mid_block.resnets.0", c, c, temb)
    if not bn:
        out += _transformer("mid_block.attentions.0", c, ctx)
    out += _resnet("mid_block.resnets.1", c, c, temb)
    # up
    for i, layers in enumerate(up_block_channels(cfg)):
        for j, (cin, _hid, _skip, cout) in enumerate(layers):
            out += _resnet(f"up_blocks.{i}.resnets.{j}", cin, cout, temb)
            if cfg.up_has_attn[i] and not bn:
                out += _transformer(f"up_blocks.{i}.attentions.{j}", cout, ctx)
        if i != len(boc) - 1:
            cout = layers[-1][3]
            out += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                    (f"up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
    if bn:
        for k, c in enumerate(tap_channels(cfg)[0]):
            out += [(f"brushnet_down_blocks.{k}.weight", (c, c, 1, 1)), (f"brushnet_down_blocks.{k}.bias", (c,))]
        c = boc[-1]
        out += [("brushnet_mid_block.weight", (c, c, 1, 1)), ("brushnet_mid_block.bias", (c,))]
        for k, c in enumerate(tap_channels(cfg)[2]):
            out += [(f"brushnet_up_blocks.{k}.weight", (c, c, 1, 1)), (f"brushnet_up_blocks.{k}.bias", (c,))]
    else:
        out += [("conv_norm_out.weight", (boc[0],)), ("conv_norm_out.bias", (boc[0],)),
                ("conv_out.weight", (cfg.out_channels, boc[0], 3, 3)), ("conv_out.bias", (cfg.out_channels,))]
    return out


def tap_channels(cfg: NetConfig):
    """Channel count of each of the BrushNet residual taps in pop order:
    (down[1 + sum(layers + has_downsampler)], mid, up[sum(layers+1 + has_upsampler)])
    (S/models/brushnet.py:325-327,356-364,369-371,441-449)."""
    boc = cfg.block_out_channels
    down = [boc[0]]
    for i, c in enumerate(boc):
        down += [c] * cfg.layers_per_block
        if i != len(boc) - 1:
            down.append(c)
    up = []
    rev = list(reversed(boc))
    for i, c in enumerate(rev):
        up += [c] * (cfg.layers_per_block + 1)
        if i != len(boc) - 1:
            up.append(c)
    return down, boc[-1], up
