"""The evaluation sweep of the reference, batched.

E/test_brushnet.py:163-267 walks the test set one sample at a time and, for each, calls the pipeline
`num_images_per_validation` (4) times with batch 1: per call PIL/numpy preprocessing on the host, one VAE encode, 50
denoise steps at net batch 2, one VAE decode.  On a B200 a net batch of 2 leaves the tensor cores mostly idle, so here the
(sample, repeat) pairs of the whole sweep are the work items: they are packed `images_per_call` at a time into ONE
pipeline pass whose every stage runs on the kernels —
    uint8 arrays -> mfb_prep_image_u8 / mfb_prep_mask_depth -> VaeEncoderEngine -> conditioning latents ->
    StepEngine.denoise (CUDA graph) -> VaeDecoderEngine -> mfb_post_image_u8 -> uint8
— and sharded over ranks by contiguous blocks of the global item index (sharding.shard_range, the rule of
`split_between_processes`, :163-168).  Each item's initial latents and VAE-sampling noise come from a generator seeded by
(seed, global item index), so an item's image does not depend on the batch it rides in or on the number of GPUs
(BASELINE config 3); the reference threads one generator through all calls, which ties results to the visiting order.

Inputs: masked RGB uint8 [S,H,W,3] and mask uint8 [S,H,W] (255 = mirror region) at the target resolution (the pipeline takes them
at their own size, E/test_brushnet.py:207-216), metric depth fp32 [S,Hd,Wd] at ANY resolution — when it differs from (H,W) it goes
through the dataset's `apply_transforms_depth` (normalise at its own resolution with `depth_mask` [S,Hd,Wd], bicubic antialiased
resize of the shorter side to the target + centre crop, E/dataset/dataset.py:98-166) fused with the nearest sampling to latent
resolution (`ops.depth_normalize` + `ops.resize_crop_bicubic(step=f)`) — and prompt embeddings.
Tokenizer / CLIP stay outside (embeddings in), as in the rest of the package."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .config import NetConfig
from .pipeline import StepEngine
from .sharding import shard_range
from .vae import VaeConfig, VaeDecoderEngine, VaeEncoderEngine

f32 = torch.float32


def item_generator(seed: int, item: int, stream: int = 0) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1_000_003 + int(item)) * 2 + int(stream))
    return g


class EvalSweep:
    def __init__(self, cfg: NetConfig, unet_sd, brushnet_sd, vae_cfg: VaeConfig, vae_sd: Dict[str, torch.Tensor],
                 scheduler_factory: Callable[[], object], H: int = 512, W: int = 512, images_per_call: int = 16,
                 repeats: int = 4, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 brushnet_conditioning_scale: float = 1.0, depth_delta: float = 0.5, device="cuda", precision: str = "bf16"):
        self.cfg, self.vae_cfg = cfg, vae_cfg
        self.dev = torch.device(device)
        self.H, self.W, self.n = H, W, images_per_call
        self.repeats, self.steps, self.guidance, self.scale, self.delta = repeats, num_inference_steps, guidance_scale, \
            brushnet_conditioning_scale, depth_delta
        self.f = 2 ** (len(vae_cfg.block_out_channels) - 1)               # VAE scale factor (8 for the SD VAE)
        self.h, self.w = H // self.f, W // self.f
        self.scheduler_factory = scheduler_factory
        n = images_per_call
        with ops.precision(precision):
            self.enc = VaeEncoderEngine(vae_cfg, vae_sd, n, H, W, self.dev)
            self.dec = VaeDecoderEngine(vae_cfg, vae_sd, n, self.h, self.w, self.dev)
        self.eng = StepEngine(cfg, unet_sd, brushnet_sd, n, self.h, self.w, self.dev, precision=precision, two_streams=True)
        z = lambda *s, dt=f32: torch.zeros(*s, device=self.dev, dtype=dt)
        self.rgb_d, self.mask_d, self.depth_d = z(n, H, W, 3, dt=torch.uint8), z(n, H, W, dt=torch.uint8), z(n, H, W)
        self.img = z(n, 3, H, W)
        self.mask_lat, self.depth_lat = z(n, 1, self.h, self.w), z(n, 1, self.h, self.w)
        self.scratch = z(n, dt=torch.int32)
        self.out_u8 = z(n, H, W, 3, dt=torch.uint8)

    def items(self, n_samples: int) -> List[Tuple[int, int]]:
        """Global work list: (sample, repeat), sample-major like the reference's nested loops."""
        return [(i, k) for i in range(n_samples) for k in range(self.repeats)]

    @torch.no_grad()
    def run(self, rgb_u8: np.ndarray, mask_u8: np.ndarray, depth: np.ndarray, prompt_embeds: torch.Tensor,
            negative_prompt_embeds: torch.Tensor, seed: int = 0, rank: int = 0, world: int = 1, depth_mask: Optional[np.ndarray] = None):
        """Returns (uint8 images [n_local, H, W, 3] as a numpy array, the list of (sample, repeat) they belong to).
        depth_mask: uint8 [S,Hd,Wd], the mask at the depth map's own resolution (needed when (Hd,Wd) != (H,W))."""
        S = rgb_u8.shape[0]
        native_depth = tuple(depth.shape[1:]) != (self.H, self.W)
        if native_depth:
            if self.H != self.W:
                raise ValueError("depth resize + centre crop produces a square map: the sweep's H and W must be equal")
            if depth_mask is None or depth_mask.shape != depth.shape:
                raise ValueError("depth at its own resolution needs `depth_mask` of the same shape (apply_transforms_depth(depth, mask))")
            Hd, Wd = depth.shape[1:]
            dsrc = torch.zeros(self.n, Hd, Wd, device=self.dev, dtype=f32)
            dmsk = torch.zeros(self.n, Hd, Wd, device=self.dev, dtype=torch.uint8)
            dnorm = torch.zeros(self.n, Hd, Wd, device=self.dev, dtype=f32)
        items = self.items(S)
        mine = [items[j] for j in shard_range(len(items), rank, world)]
        gidx = list(shard_range(len(items), rank, world))
        lc = self.cfg.in_channels
        out = np.zeros((len(mine), self.H, self.W, 3), np.uint8)
        n = self.n
        for b0 in range(0, len(mine), n):
            chunk = list(range(b0, min(b0 + n, len(mine))))
            pad = chunk + [chunk[-1]] * (n - len(chunk))                  # the last call is padded with a repeat of its last item
            si = [mine[j][0] for j in pad]
            self.rgb_d.copy_(torch.from_numpy(np.ascontiguousarray(rgb_u8[si])))
            self.mask_d.copy_(torch.from_numpy(np.ascontiguousarray(mask_u8[si])))
            if native_depth:
                dsrc.copy_(torch.from_numpy(np.ascontiguousarray(depth[si], dtype=np.float32)))
                dmsk.copy_(torch.from_numpy(np.ascontiguousarray(depth_mask[si])))
            else:
                self.depth_d.copy_(torch.from_numpy(np.ascontiguousarray(depth[si], dtype=np.float32)))
            lat0 = torch.stack([torch.randn(lc, self.h, self.w, generator=item_generator(seed, gidx[j], 0)) for j in pad])
            vnoise = torch.stack([torch.randn(self.vae_cfg.latent_channels, self.h, self.w, generator=item_generator(seed, gidx[j], 1))
                                  for j in pad])
            # preprocessing -> masked-image latents -> conditioning (pipeline_brushnet.py:1188-1202)
            ops.prep_image_u8(self.rgb_d, self.img)
            if native_depth:
                ops.prep_mask_depth(self.mask_d, None, self.mask_lat, None, None, factor=self.f)
                ops.depth_normalize(dsrc, dmsk, dnorm, self.scratch, delta=self.delta)
                ops.resize_crop_bicubic(dnorm, self.depth_lat, self.H, step=self.f)
            else:
                ops.prep_mask_depth(self.mask_d, self.depth_d, self.mask_lat, self.depth_lat, self.scratch, factor=self.f, delta=self.delta)
            lat = self.enc.encode(self.img, noise=vnoise, scale=self.vae_cfg.scaling_factor)
            cond = torch.cat([lat, self.mask_lat, self.depth_lat], 1)
            ehs = torch.cat([negative_prompt_embeds[si] if negative_prompt_embeds.shape[0] == S else negative_prompt_embeds.expand(n, -1, -1),
                             prompt_embeds[si]], 0)                      # uncond half first (:1102-1103)
            self.eng.set_conditioning(ehs.to(self.dev), torch.cat([cond, cond], 0))
            x = self.eng.denoise(lat0, self.scheduler_factory(), self.steps, self.guidance,
                                 conditioning_scales=[self.scale] * self.steps)
            img = self.dec.decode(x / self.vae_cfg.scaling_factor)       # :1337-1342
            ops.post_image_u8(img, self.out_u8)
            out[chunk[0]:chunk[-1] + 1] = self.out_u8[: len(chunk)].cpu().numpy()
        return out, mine


@torch.no_grad()
def dataset_transforms(rgb_u8: torch.Tensor, masked_rgb_u8: torch.Tensor, mask_u8: torch.Tensor, depth: Optional[torch.Tensor] = None,
                       resolution: int = 512, depth_delta: float = 0.5) -> Dict[str, torch.Tensor]:
    """The tensors `HDF5Dataset.__getitem__` builds for one training / evaluation sample (E/dataset/dataset.py:229-271), for a batch
    of equally sized samples, on the device: `apply_transforms_rgb` (image, masked image: /255, bicubic antialiased resize of the
    shorter side to `resolution`, centre crop, Normalize([0.5], [0.5]); :70-82), `apply_transforms_mask` (/255, same resize + crop;
    :84-96) and `apply_transforms_depth` (max-scene-depth normalisation over mask > 0 + delta to [-1, 1], same resize + crop; :98-166).
    rgb_u8 / masked_rgb_u8 uint8 [N,H,W,3], mask_u8 uint8 [N,H,W], depth fp32 [N,H,W] or None, all on the device.
    The normalisation 2x - 1 is applied before the resize here (the filter weights sum to 1, so it commutes up to fp32 rounding)."""
    N, Hs, Ws, _ = rgb_u8.shape
    dev = rgb_u8.device
    out: Dict[str, torch.Tensor] = {}
    tmp = torch.empty(N, 3, Hs, Ws, device=dev, dtype=f32)
    for key, src in (("pixel_values", rgb_u8), ("conditioning_pixel_values", masked_rgb_u8)):
        ops.prep_image_u8(src.contiguous(), tmp)
        o = torch.empty(N, 3, resolution, resolution, device=dev, dtype=f32)
        ops.resize_crop_bicubic(tmp, o, resolution)
        out[key] = o
    m = mask_u8.to(f32).div_(255.0).contiguous()                              # layout / dtype plumbing of the uint8 mask
    mo = torch.empty(N, 1, resolution, resolution, device=dev, dtype=f32)
    ops.resize_crop_bicubic(m, mo.view(N, resolution, resolution), resolution)
    out["masks"] = mo
    if depth is not None:
        dn = torch.empty(N, Hs, Ws, device=dev, dtype=f32)
        ops.depth_normalize(depth.contiguous(), mask_u8.contiguous(), dn, torch.zeros(N, device=dev, dtype=torch.int32), delta=depth_delta)
        do = torch.empty(N, 1, resolution, resolution, device=dev, dtype=f32)
        ops.resize_crop_bicubic(dn, do.view(N, resolution, resolution), resolution)
        out["depth"] = do
    return out
