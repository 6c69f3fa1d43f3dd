"""CPU restatement (numpy / torch) of the per-image input and output processing around the MirrorFusion pipeline —
TEST INFRASTRUCTURE ONLY (only tests/, smoke() and bench.py's cpu_baseline leg may import oracle/).

Follows, for inputs already at the target resolution: VaeImageProcessor.preprocess / postprocess
(S/image_processor.py:446-530,566-620: /255, 2x-1; (x/2+0.5).clamp(0,1), (255x).round()), the mask rule and the nearest
resampling of pipeline_brushnet.py:1139,1190-1202, and HDF5Dataset.apply_transforms_depth with
normalization_method="max_scene_depth" (E/dataset/dataset.py:98-145).  Pinned by tests/golden/prep_golden.npz, which
oracle/make_golden.py produces by calling the reference's own functions."""
from __future__ import annotations

import numpy as np


def prep_image(rgb_u8: np.ndarray) -> np.ndarray:
    """uint8 [N,H,W,3] -> fp32 [N,3,H,W] in [-1,1]."""
    x = rgb_u8.astype(np.float32) / 255.0
    return (2.0 * x - 1.0).transpose(0, 3, 1, 2)


def prep_mask(mask_u8: np.ndarray, factor: int = 8) -> np.ndarray:
    """uint8 [N,H,W] -> latent-resolution {0,1} mask [N,1,H/f,W/f]: 1 where the normalised 3-channel mask sums below 0."""
    m3 = 2.0 * (np.repeat(mask_u8[:, None].astype(np.float32), 3, 1) / 255.0) - 1.0
    full = (m3.sum(1, keepdims=True) < 0).astype(np.float32)
    return full[:, :, ::factor, ::factor]                       # F.interpolate(mode="nearest") to 1/f picks (f*i, f*j)


def prep_depth(depth: np.ndarray, mask_u8: np.ndarray, factor: int = 8, delta: float = 0.5) -> np.ndarray:
    """metric depth fp32 [N,H,W] -> [N,1,H/f,W/f] in [-1,1]."""
    out = []
    for d, m in zip(depth, mask_u8):
        dmax = d[m > 0].max() + delta
        out.append(2.0 * (np.clip(d, 0, dmax) / dmax) - 1.0)
    return np.stack(out)[:, None, ::factor, ::factor].astype(np.float32)


def post_image(img: np.ndarray) -> np.ndarray:
    """fp32 [N,3,H,W] in [-1,1] -> uint8 [N,H,W,3]."""
    x = np.clip(img / 2 + 0.5, 0, 1).transpose(0, 2, 3, 1)
    return (x * 255).round().astype(np.uint8)
