"""Generate tests/golden/train_glue.npz from the REFERENCE ITSELF (run in the build container only; needs /root/reference).

    python oracle/make_golden_train.py

Calls the reference's own `DDPMScheduler.add_noise` / `get_velocity` (S/schedulers/scheduling_ddpm.py:501-546) with the
SD1.5 training noise schedule and `diffusers.training_utils.compute_snr` (S/training_utils.py:50-73), then evaluates the
loss block of E/train_brushnet_mirror.py:1433-1450 (plain and min-SNR weighted) on seeded tensors with torch autograd.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import GOLD, import_reference  # noqa: E402
from train_oracle import resnet_block_case  # noqa: E402


def main():
    diffusers = import_reference()
    from diffusers.training_utils import compute_snr
    sched = diffusers.DDPMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    g = torch.Generator().manual_seed(4321)
    B, shape = 6, (4, 16, 16)
    x0 = torch.randn(B, *shape, generator=g)
    noise = torch.randn(B, *shape, generator=g)
    t = torch.tensor([0, 1, 17, 500, 998, 999], dtype=torch.long)
    noisy = sched.add_noise(x0, noise, t)
    vel = sched.get_velocity(x0, noise, t)
    snr = compute_snr(sched, t)

    pred = torch.randn(B, *shape, generator=g, requires_grad=True)
    out = {}
    for name, gamma in (("plain", None), ("snr5", 5.0)):
        pred.grad = None
        target = noise
        if gamma is None:
            loss = torch.nn.functional.mse_loss(pred.float(), target.float(), reduction="mean")
            w = torch.ones(B)
        else:   # train_brushnet_mirror.py:1439-1450, prediction_type == "epsilon"
            w = torch.stack([snr, gamma * torch.ones_like(t)], dim=1).min(dim=1)[0] / snr
            loss = torch.nn.functional.mse_loss(pred.float(), target.float(), reduction="none")
            loss = loss.mean(dim=list(range(1, len(loss.shape)))) * w
            loss = loss.mean()
        loss.backward()
        out[f"loss_{name}"] = loss.detach().numpy()
        out[f"grad_{name}"] = pred.grad.detach().numpy().copy()
        out[f"w_{name}"] = w.numpy()
    np.savez_compressed(os.path.join(GOLD, "train_glue.npz"), alphas_cumprod=sched.alphas_cumprod.numpy(), x0=x0.numpy(),
                        noise=noise.numpy(), t=t.numpy(), noisy=noisy.numpy(), velocity=vel.numpy(), snr=snr.numpy(),
                        pred=pred.detach().numpy(), **out)
    print("wrote train_glue.npz", {k: float(v) for k, v in out.items() if k.startswith("loss")})


def resnet_block_golden():
    """The reference's OWN ResnetBlock2D (S/models/resnet.py:184-405), forward + autograd in fp32 on the seeded case: full out / dx /
    d rowbias, and for every parameter gradient its L2 norm and 64 entries at seeded positions (keeps the fixture small)."""
    diffusers = import_reference()
    from diffusers.models.resnet import ResnetBlock2D
    out = {}
    for tag, cin, cout in (("id", 64, 64), ("sc", 64, 128)):
        sd, x, emb, d_out = resnet_block_case(cin, cout, seed=77)
        blk = ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=emb.shape[1], groups=32, eps=1e-5)
        blk.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
        keep = {}

        def hook(module, inputs, output):          # keep the projected row bias and its gradient (returns None: output unchanged)
            output.retain_grad()
            keep["rb"] = output

        blk.time_emb_proj.register_forward_hook(hook)
        x = x.clone().requires_grad_(True)
        y = blk(x, emb)
        y.backward(d_out)
        out[f"{tag}_out"], out[f"{tag}_dx"], out[f"{tag}_d_rowbias"] = y.detach().numpy(), x.grad.numpy(), keep["rb"].grad.numpy()
        gi = torch.Generator().manual_seed(5)
        for name, prm in blk.named_parameters():
            gflat = prm.grad.reshape(-1)
            idx = torch.randint(0, gflat.numel(), (64,), generator=gi)
            out[f"{tag}_g_{name}_norm"] = gflat.double().norm().numpy()
            out[f"{tag}_g_{name}_idx"] = idx.numpy()
            out[f"{tag}_g_{name}_val"] = gflat[idx].numpy()
    np.savez_compressed(os.path.join(GOLD, "resnet_block_grad.npz"), **out)
    print("wrote resnet_block_grad.npz", len(out), "arrays")


def lr_schedule_golden():
    """Multipliers of the reference's own get_scheduler (S/optimization.py:289-352) for the schedules the script can select."""
    import_reference()
    from diffusers.optimization import get_scheduler
    out = {}
    for name in ("constant", "constant_with_warmup", "linear", "cosine"):
        prm = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([prm], lr=1e-4)
        sch = get_scheduler(name, optimizer=opt, num_warmup_steps=5, num_training_steps=40)
        lrs = []
        for _ in range(45):
            lrs.append(sch.get_last_lr()[0])
            opt.step()
            sch.step()
        out[name] = np.array(lrs)
    np.savez_compressed(os.path.join(GOLD, "lr_schedules.npz"), **out)
    print("wrote lr_schedules.npz")


if __name__ == "__main__":
    main()
    resnet_block_golden()
    lr_schedule_golden()
