"""Checkpoint import / export without instantiating torch modules.

The reference stores each network as a diffusers model directory — `config.json` + `diffusion_pytorch_model.safetensors`
(`ModelMixin.save_pretrained`, S/models/modeling_utils.py:303-391) — and the fine-tuning script writes
`checkpoint-N/brushnet/` next to the frozen base `unet/` (E/train_brushnet_mirror.py:997-1032).  This module reads such a
directory straight into what the engines consume: a `NetConfig` (from `config.json`) and a flat `state_dict`
(names exactly as `UNet2DConditionModel.state_dict()` / `BrushNetModel.state_dict()`), and checks
both strictly: a config outside the SD1.5 family the kernels implement, or a tensor whose name / shape differs from the
architecture's census (`config.param_shapes`), is an error that says what differs — never a silent partial load."""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple

import torch

from .config import NetConfig, param_shapes

WEIGHTS = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin")

# config.json entries that must have exactly these values for the hot path to be the implemented one
_REQUIRED = {
    "act_fn": "silu", "flip_sin_to_cos": True, "freq_shift": 0, "resnet_time_scale_shift": "default", "only_cross_attention": False,
    "use_linear_projection": False, "upcast_attention": False, "class_embed_type": None, "addition_embed_type": None,
    "encoder_hid_dim": None, "num_class_embeds": None, "transformer_layers_per_block": 1, "mid_block_scale_factor": 1,
    "downsample_padding": 1, "dual_cross_attention": False, "time_embedding_type": "positional", "time_cond_proj_dim": None,
    "conv_in_kernel": 3, "conv_out_kernel": 3, "dropout": 0.0, "attention_type": "default",
}


class CheckpointError(ValueError):
    pass


def config_from_json(cfg: dict, net: str) -> NetConfig:
    """`config.json` of UNet2DConditionModel (net="unet") or BrushNetModel (net="brushnet") -> NetConfig."""
    want_cls = "UNet2DConditionModel" if net == "unet" else "BrushNetModel"
    if cfg.get("_class_name") != want_cls:
        raise CheckpointError(f"config.json describes {cfg.get('_class_name')!r}, expected {want_cls!r}")
    bad = {k: cfg[k] for k, v in _REQUIRED.items() if k in cfg and cfg[k] != v}
    if bad:
        raise CheckpointError(f"{want_cls} config outside the implemented SD1.5 family: {bad}")
    boc = tuple(cfg["block_out_channels"])
    n = len(boc)
    if net == "unet":
        down = tuple(t == "CrossAttnDownBlock2D" for t in cfg["down_block_types"])
        up = tuple(t == "CrossAttnUpBlock2D" for t in cfg["up_block_types"])
        known = all(t in ("CrossAttnDownBlock2D", "DownBlock2D") for t in cfg["down_block_types"]) and \
            all(t in ("CrossAttnUpBlock2D", "UpBlock2D") for t in cfg["up_block_types"]) and \
            cfg.get("mid_block_type", "UNetMidBlock2DCrossAttn") == "UNetMidBlock2DCrossAttn"
        if not known:
            raise CheckpointError(f"unsupported block types {cfg['down_block_types']} / {cfg['up_block_types']} / {cfg.get('mid_block_type')}")
    else:
        if any(t != "DownBlock2D" for t in cfg["down_block_types"]) or any(t != "UpBlock2D" for t in cfg["up_block_types"]) or \
                cfg.get("mid_block_type") != "MidBlock2D":
            raise CheckpointError("BrushNetModel with attention blocks is not the MirrorFusion branch (brushnet.py:138-214)")
        # the branch has no attention; which levels of the UNet it pairs with is the UNet's business (SD1.5 pattern)
        down = (True,) * (n - 1) + (False,)
        up = (False,) + (True,) * (n - 1)
    ahd = cfg.get("num_attention_heads") or cfg["attention_head_dim"]     # unet_2d_condition.py:231-237: "head dim" = head COUNT
    if not isinstance(ahd, int):
        raise CheckpointError(f"per-level attention_head_dim {ahd} is not supported")
    return NetConfig(in_channels=cfg["in_channels"], out_channels=cfg.get("out_channels", cfg["in_channels"]),
                     conditioning_channels=cfg.get("conditioning_channels", 6), block_out_channels=boc,
                     layers_per_block=cfg["layers_per_block"], down_has_attn=down, up_has_attn=up, heads=ahd,
                     cross_attention_dim=cfg["cross_attention_dim"], norm_num_groups=cfg["norm_num_groups"],
                     norm_eps=cfg["norm_eps"], sample_size=cfg.get("sample_size") or 64)


def _read_weights(path: str) -> Dict[str, torch.Tensor]:
    for name in WEIGHTS:
        f = os.path.join(path, name)
        if os.path.exists(f):
            if name.endswith(".safetensors"):
                from safetensors import safe_open
                out = {}
                with safe_open(f, framework="pt", device="cpu") as sf:
                    for k in sf.keys():
                        out[k] = sf.get_tensor(k)
                return out
            return torch.load(f, map_location="cpu", weights_only=True)
    raise CheckpointError(f"no {' / '.join(WEIGHTS)} in {path}")


def check_state_dict(sd: Dict[str, torch.Tensor], cfg: NetConfig, net: str):
    """Strict census check (what load_state_dict(strict=True) would do on the reference module)."""
    want = dict(param_shapes(cfg, net))
    missing = sorted(set(want) - set(sd))
    extra = sorted(set(sd) - set(want))
    wrong = sorted((k, tuple(sd[k].shape), want[k]) for k in want if k in sd and tuple(sd[k].shape) != tuple(want[k]))
    if missing or extra or wrong:
        raise CheckpointError(f"{net} checkpoint does not match the architecture: missing {missing[:5]} ({len(missing)}), "
                              f"unexpected {extra[:5]} ({len(extra)}), wrong shape {wrong[:5]} ({len(wrong)})")


def load_model_dir(path: str, net: str) -> Tuple[NetConfig, Dict[str, torch.Tensor]]:
    """One diffusers model directory (unet/ or brushnet/) -> (NetConfig, fp32 state_dict)."""
    with open(os.path.join(path, "config.json")) as f:
        cfg = config_from_json(json.load(f), net)
    sd = {k: v.float() for k, v in _read_weights(path).items()}
    check_state_dict(sd, cfg, net)
    return cfg, sd


def load_mirrorfusion(unet_dir: str, brushnet_dir: str):
    """(cfg, unet_sd, brushnet_sd) for MirrorFusionB200Pipeline / StepEngine.  The two configs must describe one family."""
    ucfg, usd = load_model_dir(unet_dir, "unet")
    bcfg, bsd = load_model_dir(brushnet_dir, "brushnet")
    for k in ("in_channels", "block_out_channels", "layers_per_block", "norm_num_groups", "norm_eps"):
        if getattr(ucfg, k) != getattr(bcfg, k):
            raise CheckpointError(f"unet and brushnet configs disagree on {k}: {getattr(ucfg, k)} vs {getattr(bcfg, k)}")
    import dataclasses
    cfg = dataclasses.replace(ucfg, conditioning_channels=bcfg.conditioning_channels)
    return cfg, usd, bsd


# ---------------------------------------------------------------------------------------------------------------- export
def brushnet_config_json(cfg: NetConfig) -> dict:
    """The `config.json` `BrushNetModel.save_pretrained` writes for the MirrorFusion branch (S/models/brushnet.py:138-214 constructor
    arguments as registered by `register_to_config`; values of the architecture-independent keys as `BrushNetModel.from_unet` sets
    them, :466-500) — what `BrushNetModel.from_pretrained(dir)` needs to rebuild the module the state_dict belongs to."""
    n = len(cfg.block_out_channels)
    return {
        "_class_name": "BrushNetModel", "_diffusers_version": "0.27.0.dev0", "act_fn": "silu", "addition_embed_type": None,
        "addition_embed_type_num_heads": 64, "addition_time_embed_dim": None, "attention_head_dim": cfg.heads,
        "block_out_channels": list(cfg.block_out_channels), "brushnet_conditioning_channel_order": "rgb", "class_embed_type": None,
        "conditioning_channels": cfg.conditioning_channels, "conditioning_embedding_out_channels": [16, 32, 96, 256],
        "cross_attention_dim": cfg.cross_attention_dim, "down_block_types": ["DownBlock2D"] * n, "downsample_padding": 1,
        "encoder_hid_dim": None, "encoder_hid_dim_type": None, "flip_sin_to_cos": True, "freq_shift": 0, "global_pool_conditions": False,
        "in_channels": cfg.in_channels, "layers_per_block": cfg.layers_per_block, "mid_block_scale_factor": 1, "mid_block_type": "MidBlock2D",
        "norm_eps": cfg.norm_eps, "norm_num_groups": cfg.norm_num_groups, "num_attention_heads": None, "num_class_embeds": None,
        "only_cross_attention": False, "projection_class_embeddings_input_dim": None, "resnet_time_scale_shift": "default",
        "transformer_layers_per_block": 1, "up_block_types": ["UpBlock2D"] * n, "upcast_attention": False, "use_linear_projection": False,
    }


def save_brushnet_dir(path: str, cfg: NetConfig, sd: Dict[str, torch.Tensor]):
    """Write a trained BrushNet as the diffusers model directory the reference's checkpoint hook writes
    (`checkpoint-N/brushnet/`: config.json + diffusion_pytorch_model.safetensors, E/train_brushnet_mirror.py:997-1032,
    S/models/modeling_utils.py:303-391), after the strict census check: `BrushNetModel.from_pretrained(path)` of the reference and
    `load_model_dir(path, "brushnet")` both read it back."""
    from safetensors.torch import save_file
    check_state_dict(sd, cfg, "brushnet")
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(brushnet_config_json(cfg), f, indent=2, sort_keys=True)
        f.write("\n")
    save_file({k: v.detach().to("cpu", torch.float32).contiguous() for k, v in sd.items()}, os.path.join(path, WEIGHTS[0]),
              metadata={"format": "pt"})
