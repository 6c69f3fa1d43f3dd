"""The error FLOOR of bf16 storage for the denoise step — TEST INFRASTRUCTURE ONLY (like everything under oracle/).

north_star's parity bar for the bf16 path is rel-L2 1e-2 on the noise prediction against the fp32 reference.  Any pipeline that keeps
weights and activations in bf16 (the reference's own `torch_dtype=torch.bfloat16` run included) pays three independent roundings
per GEMM: the weights, the GEMM inputs and the GEMM outputs.  `bf16_storage(...)` makes the fp32 oracle (oracle/mf_oracle.py, pinned to
the reference) pay exactly those and nothing else — fp32 accumulation, fp32 norms / softmax / activations — by swapping the
`torch.nn.functional` namespace the oracle calls for a proxy that rounds at the chosen points.  The resulting error is what an
IDEAL bf16-storage implementation of the same op list would show on the same weights and inputs; the GPU tests report the CUDA path's
error next to it (tests/test_gpu_model.py), and DESIGN.md §2 uses it to split the error budget (weights 5.97e-3, inputs 5.56e-3,
outputs 5.80e-3 -> 9.99e-3 on the SD1.5-shaped nets; the CUDA path: 9.73e-3, it fuses some output roundings away).
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from . import mf_oracle as O


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundingF:
    """torch.nn.functional with bf16 rounding of conv2d / linear weights, inputs and / or outputs."""

    def __init__(self, weights: bool, inputs: bool, outputs: bool):
        self.w, self.i, self.o = weights, inputs, outputs

    def __getattr__(self, name):
        return getattr(F, name)

    def conv2d(self, x, w, b=None, **kw):
        y = F.conv2d(_bf(x) if self.i else x, _bf(w) if self.w else w, b, **kw)
        return _bf(y) if self.o else y

    def linear(self, x, w, b=None):
        y = F.linear(_bf(x) if self.i else x, _bf(w) if self.w else w, b)
        return _bf(y) if self.o else y


@contextlib.contextmanager
def bf16_storage(weights: bool = True, inputs: bool = True, outputs: bool = True):
    """Within the block every conv / linear of the oracle rounds the selected operands to bf16 (accumulation stays fp32)."""
    old = O.F
    O.F = _RoundingF(weights, inputs, outputs)
    try:
        yield
    finally:
        O.F = old


@torch.no_grad()
def noise_pred_floor(unet_sd, bn_sd, cfg, latent_in, t, ehs, cond, scale: float = 1.0, ref=None, **which) -> float:
    """rel-L2 between the oracle's noise prediction with bf16 storage and its fp32 one (`ref`: the latter if already computed)."""
    if ref is None:
        ref, _ = O.noise_pred_step(unet_sd, bn_sd, cfg, latent_in, t, ehs, cond, scale)
    with bf16_storage(**which):
        got, _ = O.noise_pred_step(unet_sd, bn_sd, cfg, latent_in, t, ehs, cond, scale)
    return ((got.double() - ref.double()).norm() / ref.double().norm()).item()
