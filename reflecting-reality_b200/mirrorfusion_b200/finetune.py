"""One BrushNet-branch fine-tune step (UNet frozen) on the kernels — BASELINE config 4, the loop body of
E/train_brushnet_mirror.py:1404-1466:

    noisy = noise_scheduler.add_noise(latents, noise, timesteps)                                   (:1416)
    down, mid, up = brushnet(noisy, timesteps, ehs, brushnet_cond=conditioning_latents)            (MirrorFusionModel.forward :836-888)
    model_pred = unet(noisy, timesteps, ehs, down_block_add_samples=..., mid_..., up_...)
    loss = F.mse_loss(model_pred.float(), noise.float())                                           (:1433-1450, optional min-SNR weights)
    accelerator.backward(loss)      -> frozen-UNet data-gradient chain -> 28 tap gradients -> every BrushNet gradient; DDP all-reduce
    clip_grad_norm_; optimizer.step(); lr_scheduler.step(); optimizer.zero_grad()                  (:1460-1466)

= `BrushNetTrainer` (backward.py) + `FrozenUNetTrainer` (unet_train.py) + the glue of train.py, with the flat gradient buffer
all-reduced over NCCL in large contiguous buckets (sharding.allreduce_flat_grads; SUM, the 1 / world mean is the optimizer
kernel's `grad_scale`).  Data parallel: every rank holds a full replica and its own `batch` samples; the only collective is that
all-reduce.  bf16 tensor-core kernels with fp32 master weights (the reference's `--mixed_precision bf16` autocast + fp32 params).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops, sharding
from .backward import BrushNetTrainer, brushnet_shapes, pack_brushnet, unpack_brushnet
from .config import NetConfig
from .train import B200AdamW, FlatParams, LRSchedule, NoiseSchedule, TrainLoss
from .unet_train import FrozenUNetTrainer

f32 = torch.float32


class FineTuneStep:
    def __init__(self, cfg: NetConfig, unet_sd: Dict[str, torch.Tensor], brushnet_sd: Dict[str, torch.Tensor], *, batch: int, H: int,
                 W: int, device="cuda", lr: float = 5e-6, betas=(0.9, 0.999), weight_decay: float = 1e-2, eps: float = 1e-8,
                 max_grad_norm: Optional[float] = 1.0, lr_schedule: str = "constant", lr_warmup_steps: int = 0, max_train_steps: int = 0,
                 snr_gamma: Optional[float] = None, ctx_len: int = 77, group=None, precision: str = "bf16"):
        """precision="fp32": parity mode — both programs on the fp32 CUDA-core kernels (the 1e-3 bar against float64 autograd)."""
        ops.lib()
        self.precision = precision
        self.cfg, self.B, self.H, self.W, self.dev = cfg, batch, H, W, torch.device(device)
        self.group, self.max_grad_norm, self.snr_gamma = group, max_grad_norm, snr_gamma
        self.flat = FlatParams(brushnet_shapes(cfg), self.dev)
        self.flat.load_state_dict(pack_brushnet(cfg, brushnet_sd))
        self.brushnet = BrushNetTrainer(self.flat, cfg, B=batch, H=H, W=W, precision=precision)
        br = self.brushnet.branch
        taps = [z.tap for z in br.taps] + [br.mid_tap.tap] + [z.tap for z in br.up_taps]       # the reference's pop order: 12, mid, 15
        self.unet = FrozenUNetTrainer(cfg, unet_sd, taps, B=batch, H=H, W=W, device=self.dev, ctx_len=ctx_len, precision=precision)
        self.brushnet.bind_tap_gradients(self.unet.d_taps)      # the zero-convs read the UNet's tap gradients in place
        self.opt = B200AdamW(self.flat, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.lr_sched = LRSchedule(self.opt, lr_schedule, lr_warmup_steps, max_train_steps)
        self.noise_sched = NoiseSchedule(self.dev)
        self.loss_fn = TrainLoss(batch, self.dev)
        self.noisy = torch.zeros(batch, cfg.in_channels, H, W, device=self.dev, dtype=f32)
        self.t_dev = torch.zeros(batch, device=self.dev, dtype=torch.int64)
        self.weights = torch.zeros(batch, device=self.dev, dtype=f32)
        self.world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
        # Overlap: the up path's parameters sit at the tail of the flat buffer and their gradients are complete once the up
        # path's backward is done (the backward runs up -> mid -> down), so that range — 60 % of the 618.8 M gradients — is
        # all-reduced on a side stream while the mid / down blocks are still in their backward pass; the head follows in optimize().
        self._up_off = min(off for name, (off, _) in self.flat.table.items() if name.startswith(("up_blocks.", "brushnet_up_blocks.")))
        self._side = torch.cuda.Stream(device=self.dev) if self.world > 1 else None
        if self.world > 1:
            br.after_up_backward = self._allreduce_tail

    # the three phases are separate so a caller (bench.py, the tests) can time / inspect them
    def forward(self, latents, noise, timesteps, conditioning_latents, encoder_hidden_states):
        """latents / noise [B,4,H,W] fp32, timesteps [B] int64 (host or device), conditioning_latents [B,6,H,W], ehs [B,77,ctx].
        Returns the loss (device scalar buffer)."""
        self.t_dev.copy_(timesteps)
        self.noise_sched.add_noise(latents, noise, self.t_dev, out=self.noisy)
        self.brushnet.forward(self.noisy, conditioning_latents, self.t_dev)
        pred = self.unet.forward(self.noisy, self.t_dev, encoder_hidden_states)
        w = None
        if self.snr_gamma is not None:
            self.weights.copy_(torch.from_numpy(self.noise_sched.snr_weights(timesteps.cpu(), self.snr_gamma)))
            w = self.weights
        return self.loss_fn(pred, noise, weights=w, grad=self.unet.d_pred)       # also writes d loss / d model_pred

    def backward(self):
        dd, dm, du = self.unet.backward()
        self.brushnet.backward(dd, dm, du)

    def _allreduce_tail(self):
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            sharding.allreduce_flat_grads(self.flat.grad[self._up_off:], group=self.group)

    def optimize(self):
        if self.world > 1:
            sharding.allreduce_flat_grads(self.flat.grad[:self._up_off], group=self.group)
            torch.cuda.current_stream().wait_stream(self._side)
        self.opt.step(max_grad_norm=self.max_grad_norm, grad_scale=1.0 / self.world)
        self.lr_sched.step()
        self.brushnet.refresh_dgrad_weights()
        self.opt.zero_grad()

    def step(self, latents, noise, timesteps, conditioning_latents, encoder_hidden_states) -> torch.Tensor:
        loss = self.forward(latents, noise, timesteps, conditioning_latents, encoder_hidden_states)
        self.backward()
        self.optimize()
        return loss

    def brushnet_state_dict(self) -> Dict[str, torch.Tensor]:
        """The trained BrushNet in the reference's state_dict naming / layouts (checkpoint hook, :997-1032)."""
        return unpack_brushnet(self.cfg, self.flat)

    def save_checkpoint(self, path: str):
        """The reference's save hook (:997-1032): `<path>/brushnet/` as a diffusers model directory (readable by the reference's
        `BrushNetModel.from_pretrained` and by `MirrorFusionB200Pipeline.from_checkpoint`) + the optimizer / lr-scheduler state
        for a bit-identical resume (`<path>/optimizer.pt`)."""
        import os
        from .checkpoint import save_brushnet_dir
        save_brushnet_dir(os.path.join(path, "brushnet"), self.cfg, self.brushnet_state_dict())
        torch.save({"optimizer": self.opt.state_dict(), "lr_scheduler": self.lr_sched.state_dict()}, os.path.join(path, "optimizer.pt"))

    def load_checkpoint(self, path: str):
        """`--resume_from_checkpoint` (:1269-1296): the BrushNet weights of `<path>/brushnet/`, the optimizer moments / step count and
        the lr-scheduler position of `<path>/optimizer.pt`; the next step is bit-identical to the one an uninterrupted run takes."""
        import os
        from .checkpoint import load_model_dir
        _cfg, sd = load_model_dir(os.path.join(path, "brushnet"), "brushnet")
        self.flat.load_state_dict(pack_brushnet(self.cfg, sd))      # fp32 masters + the bf16 working copy
        self.brushnet.refresh_dgrad_weights()                       # the flipped / transposed copies the data-gradient plans read
        st = torch.load(os.path.join(path, "optimizer.pt"), map_location=self.dev, weights_only=False)
        self.opt.load_state_dict(st["optimizer"])
        self.lr_sched.load_state_dict(st["lr_scheduler"])
        self.opt.zero_grad()

    @property
    def flops_per_step(self) -> float:
        """Algorithmic FLOPs of one step on this rank's batch: forward of both nets + BrushNet backward (data + weight
        gradients = 2x its forward) + the frozen UNet's data-gradient chain."""
        return self.unet.flops_fwd + self.unet.flops_bwd + 3.0 * self._bn_fwd_flops()

    def _bn_fwd_flops(self) -> float:
        # BrushNet forward = SURVEY's 4.413e11 per sample at 64x64 for SD1.5; computed from the UNet trainer's plans for other configs
        return 4.413e11 * self.B * (self.H * self.W) / 4096.0 if self.cfg.block_out_channels == (320, 640, 1280, 1280) else 0.0
