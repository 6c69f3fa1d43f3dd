"""CPU restatement of the dataset's resize — TEST INFRASTRUCTURE ONLY (like everything under oracle/).

`transforms.Resize(res, interpolation=BICUBIC)` + `transforms.CenterCrop(res)` on float tensors, as
E/dataset/dataset.py:70-76 (RGB), :86-92 (mask) and :155-165 (normalised depth) apply them.  On tensors torchvision
(functional `resize`, antialias=True) calls F.interpolate(mode="bicubic", align_corners=False, antialias=True): ATen's separable
`_upsample_bicubic2d_aa` (Keys cubic a = -0.5, support scaled by the down-scale factor, weights normalised per output index),
horizontal pass first.  Pinned against torchvision itself in tests/test_oracle_golden.py::test_resize_oracle_vs_torchvision.
"""
from __future__ import annotations

import numpy as np


def _cubic(x, a=-0.5):
    x = np.abs(x)
    return np.where(x < 1, ((a + 2) * x - (a + 3)) * x * x + 1, np.where(x < 2, (((x - 5) * x + 8) * x - 4) * a, 0.0))


def _weights(in_size: int, out_size: int):
    """[(first input index, normalised weights)] per output index (UpSampleKernel.cpp `_compute_indices_min_size_weights_aa`)."""
    scale = np.float32(in_size) / np.float32(out_size)
    support = np.float32(2.0) * scale if scale >= 1 else np.float32(2.0)
    inv = np.float32(1.0) / scale if scale >= 1 else np.float32(1.0)
    out = []
    for o in range(out_size):
        center = scale * np.float32(o + 0.5)
        lo = max(int(center - support + np.float32(0.5)), 0)
        n = min(int(center + support + np.float32(0.5)), in_size) - lo
        w = _cubic((np.arange(n, dtype=np.float32) + np.float32(lo) - center + np.float32(0.5)) * inv).astype(np.float32)
        out.append((lo, w / w.sum(dtype=np.float32)))
    return out


def resized_size(Hs: int, Ws: int, res: int):
    """torchvision `_compute_resized_output_size` for an int size: shorter side -> res."""
    return (res, int(res * Ws / Hs)) if Hs <= Ws else (int(res * Hs / Ws), res)


def crop_offsets(Hr: int, Wr: int, res: int):
    """torchvision `center_crop`: int(round((size - crop) / 2.0)) (Python round: half to even)."""
    return int(round((Hr - res) / 2.0)), int(round((Wr - res) / 2.0))


def resize_crop_bicubic(x: np.ndarray, res: int, step: int = 1) -> np.ndarray:
    """x [..., Hs, Ws] float32 -> [..., res/step, res/step]: Resize(res, BICUBIC, antialias) + CenterCrop(res), every `step`-th pixel."""
    x = np.asarray(x, np.float32)
    Hs, Ws = x.shape[-2:]
    Hr, Wr = resized_size(Hs, Ws, res)
    top, left = crop_offsets(Hr, Wr, res)
    wy, wx = _weights(Hs, Hr), _weights(Ws, Wr)
    n = res // step
    hor = np.zeros(x.shape[:-1] + (n,), np.float32)                       # horizontal pass
    for j in range(n):
        lo, w = wx[left + j * step]
        hor[..., j] = (x[..., lo:lo + len(w)] * w).sum(-1, dtype=np.float32)
    out = np.zeros(x.shape[:-2] + (n, n), np.float32)                     # vertical pass
    for i in range(n):
        lo, w = wy[top + i * step]
        out[..., i, :] = (hor[..., lo:lo + len(w), :] * w[:, None]).sum(-2, dtype=np.float32)
    return out
