"""CPU restatement (plain PyTorch fp32) of AutoencoderKL.decode — TEST INFRASTRUCTURE ONLY (see mf_oracle.py's header:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import anything under oracle/).

Follows S/models/autoencoders/autoencoder_kl.py:280-309 (post_quant_conv -> Decoder), S/models/autoencoders/vae.py:284-349
(Decoder.forward), S/models/unets/unet_2d_blocks.py:748-770 (UNetMidBlock2D.forward), :2832-2842 (UpDecoderBlock2D.forward),
S/models/resnet.py:329-405 (ResnetBlock2D without temb), S/models/attention_processor.py:1204-1286 (single head,
group_norm, biased q/k/v, residual_connection, rescale_output_factor 1), S/models/upsampling.py:145-186.
Pinned against the reference's own output by tests/test_oracle_golden.py (tests/golden/tiny_vae_decode.npz)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _resnet(sd, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    h = F.silu(F.group_norm(x, groups, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps))
    h = F.conv2d(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, groups, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps))
    h = F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def _attention(sd, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    b, c, h, w = x.shape
    t = x.view(b, c, h * w)
    t = F.group_norm(t, groups, sd[f"{p}.group_norm.weight"], sd[f"{p}.group_norm.bias"], eps).transpose(1, 2)
    q = F.linear(t, sd[f"{p}.to_q.weight"], sd[f"{p}.to_q.bias"])
    k = F.linear(t, sd[f"{p}.to_k.weight"], sd[f"{p}.to_k.bias"])
    v = F.linear(t, sd[f"{p}.to_v.weight"], sd[f"{p}.to_v.bias"])
    s = torch.softmax(q @ k.transpose(-1, -2) * (c ** -0.5), dim=-1)          # one head of dim c
    o = F.linear(s @ v, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])
    return o.transpose(1, 2).reshape(b, c, h, w) + x


@torch.no_grad()
def vae_decode(sd: Dict[str, Tensor], cfg, z: Tensor) -> Tensor:
    """sd: AutoencoderKL.state_dict() keys (post_quant_conv.*, decoder.*); cfg: mirrorfusion_b200.vae.VaeConfig."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    d = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    x = F.conv2d(z, d["conv_in.weight"], d["conv_in.bias"], padding=1)
    x = _resnet(d, "mid_block.resnets.0", x, g, eps)
    x = _attention(d, "mid_block.attentions.0", x, g, eps)
    x = _resnet(d, "mid_block.resnets.1", x, g, eps)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = _resnet(d, f"up_blocks.{i}.resnets.{j}", x, g, eps)
        if i != n - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, d[f"up_blocks.{i}.upsamplers.0.conv.weight"], d[f"up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(F.group_norm(x, g, d["conv_norm_out.weight"], d["conv_norm_out.bias"], eps))
    return F.conv2d(x, d["conv_out.weight"], d["conv_out.bias"], padding=1)


@torch.no_grad()
def vae_encode_moments(sd: Dict[str, Tensor], cfg, x: Tensor):
    """AutoencoderKL.encode(x).latent_dist as (mean, logvar): S/models/autoencoders/autoencoder_kl.py:238-268,
    Encoder.forward (S/models/autoencoders/vae.py:139-176), DownEncoderBlock2D with Downsample2D(padding=0)
    (S/models/unets/unet_2d_blocks.py:1564-1575, S/models/downsampling.py:141-143: F.pad (0,1,0,1) then conv stride 2),
    DiagonalGaussianDistribution (vae.py:769-776: chunk, logvar clamped to [-30, 20])."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    e = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    h = F.conv2d(x, e["conv_in.weight"], e["conv_in.bias"], padding=1)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = _resnet(e, f"down_blocks.{i}.resnets.{j}", h, g, eps)
        if i != n - 1:
            h = F.pad(h, (0, 1, 0, 1))
            h = F.conv2d(h, e[f"down_blocks.{i}.downsamplers.0.conv.weight"], e[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    h = _resnet(e, "mid_block.resnets.0", h, g, eps)
    h = _attention(e, "mid_block.attentions.0", h, g, eps)
    h = _resnet(e, "mid_block.resnets.1", h, g, eps)
    h = F.silu(F.group_norm(h, g, e["conv_norm_out.weight"], e["conv_norm_out.bias"], eps))
    h = F.conv2d(h, e["conv_out.weight"], e["conv_out.bias"], padding=1)
    m = F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])
    mean, logvar = m.chunk(2, dim=1)
    return mean, logvar.clamp(-30.0, 20.0)


def latent_sample(mean: Tensor, logvar: Tensor, noise: Tensor) -> Tensor:
    """DiagonalGaussianDistribution.sample with the noise given (vae.py:782-791)."""
    return mean + torch.exp(0.5 * logvar) * noise
