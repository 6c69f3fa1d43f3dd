"""The UNMODIFIED reference, run through its own public API — the baseline leg of bench.py (never part of `value` / `e2e`).

`baseline/_ref/` is a plain `pip install --target` of the reference package (the diffusers 0.27 fork under
/root/reference/MirrorFusion; recipe: `baseline/install_ref.sh`, also run by `__graft_entry__.build()` where the reference is
mounted).  It is git-ignored but NOT gpurun-ignored, so it travels to the GPU box with the snapshot.  This module only
 * finds it (`baseline/_ref` first, the read-only mount second),
 * builds a `StableDiffusionBrushNetPipeline` (pipeline_brushnet.py:185-199) around the reference's own `UNet2DConditionModel`,
   `BrushNetModel.from_unet`, `AutoencoderKL`, `UniPCMultistepScheduler.from_config(DDIMScheduler.config)` (E/test_brushnet.py:158)
   with the seeded SD1.5-shaped random-init weights of `mirrorfusion_b200.synth` (strict `load_state_dict`), and
 * calls `pipe(...)` (`__call__`, :848-880) with SynMirror-shaped synthetic tensors, timing every denoise step from
   `callback_on_step_end` (:1317-1325): host clock on the CPU, CUDA events on a GPU.
Nothing of mirrorfusion_b200's engine, kernels or oracle is on that path (only its seeded weight / input generators).
"""
from __future__ import annotations

import os
import statistics
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_INSTALL = os.path.join(HERE, "_ref")
REF_MOUNT = "/root/reference/MirrorFusion/src"


def find_reference():
    """-> (path, kind) of an importable reference tree, or (None, why)."""
    if os.path.isdir(os.path.join(REF_INSTALL, "diffusers")):
        return REF_INSTALL, "baseline/_ref (pip install --target of the reference)"
    if os.path.isdir(os.path.join(REF_MOUNT, "diffusers")):
        return REF_MOUNT, "/root/reference mount"
    return None, "reference not installed under baseline/_ref and /root/reference not mounted"


def import_reference():
    path, kind = find_reference()
    if path is None:
        raise ImportError(kind)
    import transformers.utils as tu       # shim in OUR harness: symbol removed in transformers 5, imported by pipeline_loading_utils.py:44
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    if path not in sys.path:
        sys.path.insert(0, path)
    import diffusers
    if not os.path.abspath(diffusers.__file__).startswith(os.path.abspath(path)):
        raise ImportError(f"another diffusers is already imported: {diffusers.__file__}")
    return diffusers, kind


def build_pipeline(diffusers, cfg, device, dtype, seed: int = 0):
    from mirrorfusion_b200.synth import make_state_dict
    down = tuple("CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg.down_has_attn)
    up = tuple("CrossAttnUpBlock2D" if a else "UpBlock2D" for a in cfg.up_has_attn)
    unet = diffusers.UNet2DConditionModel(
        sample_size=cfg.sample_size, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
        block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block, down_block_types=down, up_block_types=up,
        cross_attention_dim=cfg.cross_attention_dim, attention_head_dim=cfg.heads, norm_num_groups=cfg.norm_num_groups).eval()
    bn = diffusers.BrushNetModel.from_unet(unet, conditioning_channels=cfg.conditioning_channels).eval()
    # from_unet aliases conv_in_condition.bias to unet.conv_in.bias (brushnet.py:518): un-share before loading two state dicts
    bn.conv_in_condition.bias = torch.nn.Parameter(bn.conv_in_condition.bias.detach().clone())
    unet.load_state_dict(make_state_dict(cfg, "unet", seed), strict=True)
    bn.load_state_dict(make_state_dict(cfg, "brushnet", seed), strict=True)
    torch.manual_seed(seed)
    if len(cfg.block_out_channels) == 4 and cfg.block_out_channels[0] == 320:       # SD-VAE shape
        vae = diffusers.AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                                      up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
                                      layers_per_block=2, latent_channels=4, scaling_factor=0.18215).eval()
    else:
        vae = diffusers.AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                                      up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(16, 16, 32, 32), layers_per_block=1,
                                      latent_channels=4, norm_num_groups=8, scaling_factor=0.18215).eval()
    base = diffusers.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                                   set_alpha_to_one=False, steps_offset=1)
    sched = diffusers.UniPCMultistepScheduler.from_config(base.config)                # E/test_brushnet.py:158
    pipe = diffusers.StableDiffusionBrushNetPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, brushnet=bn, scheduler=sched,
                                                     safety_checker=None, feature_extractor=None, requires_safety_checker=False,
                                                     depth_conditioning_mode="concat")          # MirrorFusion: depth concatenated (:1196-1200)
    pipe.set_progress_bar_config(disable=True)
    return pipe.to(device=device, dtype=dtype)


def synth_pixels(images: int, side: int, seed: int = 1234):
    """RGB in [0,1], 3-channel {0,1} rectangle mask covering 20-40 % of the frame, depth in [-1,1] (SURVEY.md §8d inputs)."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(images, 3, side, side, generator=g)
    mask = torch.zeros(images, 3, side, side)
    for i in range(images):
        fh = int(side * (0.45 + 0.15 * torch.rand((), generator=g).item()))
        fw = int(side * (0.45 + 0.15 * torch.rand((), generator=g).item()))
        y0 = int((side - fh) * torch.rand((), generator=g).item())
        x0 = int((side - fw) * torch.rand((), generator=g).item())
        mask[i, :, y0:y0 + fh, x0:x0 + fw] = 1.0
    yy = torch.linspace(-1, 1, side).view(1, 1, side, 1)
    xx = torch.linspace(-1, 1, side).view(1, 1, 1, side)
    a = torch.rand(images, 1, 1, 1, generator=g) - 0.5
    b = torch.rand(images, 1, 1, 1, generator=g) - 0.5
    depth = (a * yy + b * xx + 0.3 * torch.exp(-4 * (yy * yy + xx * xx))).clamp(-1, 1)
    return rgb * (1 - mask), mask, depth


@torch.no_grad()
def run(cfg, images: int, latent: int, steps: int, warmup: int, device: str = "cpu", dtype=torch.float32, guidance: float = 7.5,
        threads: int | None = None):
    """One `pipe(...)` call of warmup + steps UniPC steps on `images` images; -> dict(sec_per_step (median of the timed steps), …).
    `output_type="latent"`: the loop is what is timed (the VAE decode is outside BASELINE.json's per-step metric)."""
    diffusers, where = import_reference()
    if device == "cpu":
        torch.set_num_threads(threads or os.cpu_count() or 1)
    pipe = build_pipeline(diffusers, cfg, device, dtype)
    side = latent * 8
    masked, mask, depth = synth_pixels(images, side)
    g = torch.Generator().manual_seed(7)
    pe = torch.randn(images, 77, cfg.cross_attention_dim, generator=g).to(device, dtype)
    ne = torch.randn(images, 77, cfg.cross_attention_dim, generator=g).to(device, dtype)
    lat = torch.randn(images, 4, latent, latent, generator=g).to(device, dtype)
    cuda = device != "cpu"
    marks = []

    def mark():
        if cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)
        else:
            marks.append(time.perf_counter())

    def on_step_end(p, i, t, kw):      # pipeline_brushnet.py:1317-1325
        mark()
        return {}

    orig_prepare = pipe.prepare_extra_step_kwargs

    def prepare_and_mark(*a, **k):      # last call before the loop (:1222): the loop's start mark
        r = orig_prepare(*a, **k)
        mark()
        return r

    pipe.prepare_extra_step_kwargs = prepare_and_mark
    n = warmup + steps
    out = pipe(prompt=None, image=masked.to(device, dtype), mask=mask.to(device, dtype), depth=depth.to(device, dtype),
               num_inference_steps=n, guidance_scale=guidance, prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat,
               brushnet_conditioning_scale=1.0, output_type="latent", return_dict=False, callback_on_step_end=on_step_end)
    if cuda:
        torch.cuda.synchronize()
        dts = [marks[i].elapsed_time(marks[i + 1]) * 1e-3 for i in range(n)]
    else:
        dts = [marks[i + 1] - marks[i] for i in range(n)]
    timed = dts[warmup:]
    final = out[0]
    return {"sec_per_step": statistics.median(timed), "sec_per_step_mean": sum(timed) / len(timed), "steps_timed": len(timed),
            "images": images, "where": where, "diffusers": getattr(diffusers, "__version__", "?"),
            "latents_finite": bool(torch.isfinite(final.float()).all()), "latents_shape": list(final.shape),
            "threads": torch.get_num_threads() if not cuda else None}
