"""One fine-tune step (config 4) for ncu's launch list: python tools/profile_train_step.py --batch 8 [--steps 1]."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
from mirrorfusion_b200.config import SD15
from mirrorfusion_b200.finetune import FineTuneStep
from mirrorfusion_b200.synth import make_inputs, make_state_dict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    cfg, B, H = SD15, a.batch, 64
    ft = FineTuneStep(cfg, make_state_dict(cfg, "unet"), make_state_dict(cfg, "brushnet"), batch=B, H=H, W=H)
    g = torch.Generator().manual_seed(1)
    inp = make_inputs(cfg, B, seed=7, cfg_duplicate=False)
    lat, noise = torch.randn(B, 4, H, H, generator=g).cuda(), torch.randn(B, 4, H, H, generator=g).cuda()
    t = torch.randint(0, 1000, (B,), generator=g)
    cond, ehs = inp["conditioning_latents"].cuda(), inp["prompt_embeds"].cuda()
    for _ in range(a.steps + 1):
        ft.step(lat, noise, t, cond, ehs)
    torch.cuda.synchronize()
    print("MARK: last step done")


if __name__ == "__main__":
    main()
