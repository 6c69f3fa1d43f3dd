"""Host side of the BrushNet fine-tune step's glue (BASELINE config 4), over the kernels of
csrc/train.cu.  It mirrors the objects E/train_brushnet_mirror.py uses around the two nets:

    noise_scheduler.add_noise(latents, noise, timesteps)   -> NoiseSchedule.add_noise        (:1404-1416)
    F.mse_loss / min-SNR weighting                         -> TrainLoss                      (:1433-1450)
    accelerator.clip_grad_norm_ + torch.optim.AdamW.step   -> B200AdamW.step(max_grad_norm)  (:1460-1466)
    DDP gradient all-reduce                                 -> sharding.allreduce_flat_grads

Layout: every trainable tensor is a view into ONE flat fp32 master buffer (`FlatParams`), with flat gradient and moment
buffers beside it and a flat bf16 working copy the tensor-core kernels read; clip + AdamW + re-quantisation is one kernel
launch and the gradient all-reduce a few large contiguous NCCL calls.  The backward of the two nets themselves is NOT here
(DESIGN.md §8): what exists of it are the conv / linear data gradient (`ops.pack_conv_dgrad_weight` + the igemm plan) and
weight gradient (`ops.conv_wgrad`).  No torch math on the path; there is no CPU fallback."""
from __future__ import annotations

import math
from typing import Dict, Iterable, Mapping, Optional, Tuple

import numpy as np
import torch

from . import ops
from .schedulers import _alphas_cumprod

_ALIGN = 8   # elements: every tensor starts on a 16-byte boundary of the bf16 working copy (TMA tensor-map bases) and 32 B of the fp32 buffers


def flat_layout(shapes: Mapping[str, Tuple[int, ...]]) -> Tuple[Dict[str, Tuple[int, int]], int]:
    """name -> (offset, numel) in the flat buffers, in the mapping's order, each offset a multiple of 8 elements."""
    table, off = {}, 0
    for name, shp in shapes.items():
        n = int(np.prod(shp)) if len(shp) else 1
        table[name] = (off, n)
        off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
    return table, off


class FlatParams:
    """Flat fp32 master parameters + gradients + AdamW moments (+ bf16 working copy) with named views."""

    def __init__(self, shapes: Mapping[str, Tuple[int, ...]], device, with_bf16: bool = True):
        self.shapes = {k: tuple(v) for k, v in shapes.items()}
        self.table, self.numel = flat_layout(self.shapes)
        z = lambda dt: torch.zeros(max(self.numel, _ALIGN), device=device, dtype=dt)
        self.param, self.grad, self.exp_avg, self.exp_avg_sq = z(torch.float32), z(torch.float32), z(torch.float32), z(torch.float32)
        self.work = z(torch.bfloat16) if with_bf16 else None

    @classmethod
    def from_state_dict(cls, sd: Mapping[str, torch.Tensor], device, names: Optional[Iterable[str]] = None, with_bf16=True):
        names = list(sd.keys()) if names is None else list(names)
        self = cls({k: tuple(sd[k].shape) for k in names}, device, with_bf16)
        self.load_state_dict({k: sd[k] for k in names})
        return self

    def _view(self, buf, name):
        off, n = self.table[name]
        return buf[off:off + n].view(self.shapes[name])

    def p(self, name): return self._view(self.param, name)
    def g(self, name): return self._view(self.grad, name)
    def w(self, name): return self._view(self.work, name)

    def load_state_dict(self, sd: Mapping[str, torch.Tensor]):
        for k, t in sd.items():
            self.p(k).copy_(t.to(torch.float32))
        if self.work is not None:
            ops.f32_to_bf16(self.param, self.work)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self.p(k).clone() for k in self.table}


class B200AdamW:
    """torch.optim.AdamW over a FlatParams (same hyper-parameters, same update, `param_groups[0]["lr"]` writable so an
    lr scheduler can drive it), fused with clip_grad_norm_ and the bf16 re-quantisation.  `grad_scale` multiplies the
    gradient first (1 / world size after a SUM all-reduce, 1 / gradient_accumulation_steps)."""

    _RING = 8
    _NH = 12     # MFB_ADAMW_HYPER_FLOATS

    def __init__(self, flat: FlatParams, lr=5e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.flat = flat
        self.param_groups = [dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        self.step_count = 0
        dev = flat.param.device
        self._hyper = torch.zeros(self._NH, device=dev, dtype=torch.float32)
        # pinned staging ring for the per-step scalars: a slot is rewritten only after the async copy that read it has
        # finished (the host never waits on the GPU unless it runs more than _RING optimizer steps ahead)
        self._ring = torch.zeros(self._RING, self._NH, dtype=torch.float32)
        self._ring_ev = [None] * self._RING
        if dev.type == "cuda":
            self._ring = self._ring.pin_memory()
        self._sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self._ws = torch.zeros(ops.SQNORM_WS_FLOATS, device=dev, dtype=torch.float32)

    def hyper(self, step: int, grad_scale: float = 1.0) -> np.ndarray:
        """The 12 scalars of mfb_adamw_step for optimizer step `step` (1-based), evaluated in float64 on the host (the
        derived ones — 1-beta1, 1-beta2, 1-lr*wd, lr/bc1 — exactly as torch/optim/adamw.py evaluates them in Python)."""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        bc1 = 1.0 - b1 ** step
        return np.array([g["lr"], b1, b2, g["eps"], g["weight_decay"], bc1, math.sqrt(1.0 - b2 ** step), grad_scale,
                         1.0 - b1, 1.0 - b2, 1.0 - g["lr"] * g["weight_decay"], g["lr"] / bc1], dtype=np.float64)

    def step(self, max_grad_norm: Optional[float] = None, grad_scale: float = 1.0):
        self.step_count += 1
        slot = self.step_count % self._RING
        if self._ring_ev[slot] is not None:
            self._ring_ev[slot].synchronize()
        self._ring[slot].copy_(torch.from_numpy(self.hyper(self.step_count, grad_scale).astype(np.float32)))
        self._hyper.copy_(self._ring[slot], non_blocking=True)
        if self._hyper.is_cuda:
            self._ring_ev[slot] = torch.cuda.Event()
            self._ring_ev[slot].record()
        f = self.flat
        sq = None
        if max_grad_norm is not None and max_grad_norm > 0:
            ops.grad_sqnorm(f.grad, self._ws, self._sq)
            sq = self._sq
        ops.adamw_step(f.param, f.grad, f.exp_avg, f.exp_avg_sq, self._hyper, param_bf16=f.work, grad_sqnorm=sq,
                       max_grad_norm=max_grad_norm or 0.0)

    def grad_norm(self, grad_scale: float = 1.0) -> float:
        """total_norm of the last clipped step (what clip_grad_norm_ returns); a device->host read."""
        return float(self._sq.sqrt().item()) * abs(grad_scale)

    def zero_grad(self, set_to_none: bool = False):
        self.flat.grad.zero_()

    def state_dict(self) -> Dict[str, object]:
        """What `accelerator.save_state` keeps of torch.optim.AdamW (E/train_brushnet_mirror.py:1488-1509): the two moment
        buffers (flat, host copies), the step count that drives the bias correction, and the hyper-parameters."""
        return {"exp_avg": self.flat.exp_avg.detach().cpu().clone(), "exp_avg_sq": self.flat.exp_avg_sq.detach().cpu().clone(),
                "step_count": int(self.step_count), "param_groups": [dict(g) for g in self.param_groups],
                "layout": dict(self.flat.table)}

    def load_state_dict(self, sd: Mapping[str, object]):
        """Resume (`--resume_from_checkpoint`, :1271-1300): save -> load -> step is bit-identical to an uninterrupted run."""
        if dict(sd["layout"]) != dict(self.flat.table):
            raise ValueError("optimizer state was saved for a different flat parameter layout")
        self.flat.exp_avg.copy_(sd["exp_avg"])
        self.flat.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step_count"])
        self.param_groups = [dict(g, betas=tuple(g["betas"])) for g in sd["param_groups"]]


class NoiseSchedule:
    """The training-time half of DDPMScheduler (SD1.5 schedule by default): add_noise, get_velocity, min-SNR weights."""

    def __init__(self, device, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="epsilon"):
        self.acp_host = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas_cumprod = torch.from_numpy(self.acp_host).to(device)
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type

    def sample_timesteps(self, B: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """Host int64 [B], uniform over the training timesteps (train_brushnet_mirror.py:1409-1410)."""
        return torch.randint(0, self.num_train_timesteps, (B,), generator=generator, dtype=torch.int64)

    def add_noise(self, x0, noise, timesteps_dev, out=None):
        out = torch.empty_like(x0) if out is None else out
        ops.add_noise(x0, noise, timesteps_dev, self.alphas_cumprod, noisy=out)
        return out

    def get_velocity(self, x0, noise, timesteps_dev, out=None):
        out = torch.empty_like(x0) if out is None else out
        ops.add_noise(x0, noise, timesteps_dev, self.alphas_cumprod, velocity=out)
        return out

    def snr_weights(self, timesteps_host: torch.Tensor, snr_gamma: float) -> np.ndarray:
        """min(SNR, gamma) / SNR (epsilon) or / (SNR + 1) (v_prediction), fp32 [B], on the host (B scalars)."""
        t = timesteps_host.numpy()
        a = self.acp_host.astype(np.float32)
        snr = ((a ** 0.5)[t] / ((1.0 - a) ** 0.5)[t]) ** 2
        w = np.minimum(snr, np.float32(snr_gamma))
        if self.prediction_type == "epsilon":
            return (w / snr).astype(np.float32)
        if self.prediction_type == "v_prediction":
            return (w / (snr + 1)).astype(np.float32)
        raise ValueError(f"Unknown prediction type {self.prediction_type}")


class TrainLoss:
    """F.mse_loss(model_pred.float(), target.float()) (optionally min-SNR weighted) and d loss / d model_pred."""

    def __init__(self, B: int, device):
        self.ws = torch.zeros(B * ops.MSE_MAX_CHUNKS, device=device, dtype=torch.float32)
        self.loss = torch.zeros(1, device=device, dtype=torch.float32)
        self.per_sample = torch.zeros(B, device=device, dtype=torch.float32)

    def __call__(self, pred, target, weights=None, grad=None):
        ops.mse_loss(pred, target, self.loss, self.ws, weights=weights, per_sample=self.per_sample, grad=grad)
        return self.loss


class LRSchedule:
    """The learning-rate schedules `diffusers.optimization.get_scheduler` builds for the fine-tune script
    (S/optimization.py:40-78,123-185; E/train_brushnet_mirror.py:1257-1265; default "constant"): a multiplier of the initial
    learning rate as a function of the optimizer step, applied by writing `param_groups[0]["lr"]` like torch's LambdaLR does.
    `step()` after every optimizer step; `get_last_lr()` for logging (:1517)."""

    NAMES = ("constant", "constant_with_warmup", "linear", "cosine")

    def __init__(self, optimizer: B200AdamW, name: str = "constant", num_warmup_steps: int = 0, num_training_steps: int = 0,
                 num_cycles: float = 0.5):
        if name not in self.NAMES:
            raise ValueError(f"unknown lr schedule {name!r}; supported: {self.NAMES}")
        self.opt, self.name, self.warmup, self.total, self.cycles = optimizer, name, num_warmup_steps, num_training_steps, num_cycles
        self.base_lr = optimizer.param_groups[0]["lr"]
        self.last_epoch = 0
        self._apply()

    def multiplier(self, step: int) -> float:
        if self.name == "constant":
            return 1.0
        if step < self.warmup:
            return float(step) / float(max(1, self.warmup))
        if self.name == "constant_with_warmup":
            return 1.0
        if self.name == "linear":
            return max(0.0, float(self.total - step) / float(max(1, self.total - self.warmup)))
        progress = float(step - self.warmup) / float(max(1, self.total - self.warmup))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(self.cycles) * 2.0 * progress)))

    def _apply(self):
        self.opt.param_groups[0]["lr"] = self.base_lr * self.multiplier(self.last_epoch)

    def step(self):
        self.last_epoch += 1
        self._apply()

    def get_last_lr(self):
        return [self.opt.param_groups[0]["lr"]]

    def state_dict(self) -> Dict[str, object]:
        return {"name": self.name, "last_epoch": int(self.last_epoch), "base_lr": float(self.base_lr), "num_warmup_steps": self.warmup,
                "num_training_steps": self.total, "num_cycles": self.cycles}

    def load_state_dict(self, sd: Mapping[str, object]):
        """Restores `base_lr` too: a scheduler constructed over a resumed optimizer would otherwise take the already-decayed
        learning rate as its base."""
        if sd["name"] != self.name:
            raise ValueError(f"lr schedule state is for {sd['name']!r}, this schedule is {self.name!r}")
        self.last_epoch, self.base_lr = int(sd["last_epoch"]), float(sd["base_lr"])
        self.warmup, self.total, self.cycles = sd["num_warmup_steps"], sd["num_training_steps"], sd["num_cycles"]
        self._apply()
