"""Is the sweep's output a pure function of (seed, global item index)?  SD1.5 geometry, few steps.  Compares, item by item:
run A (world 1) vs run A again (run-to-run), vs the two shards of a simulated world of 2 on the same engine."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import numpy as np
import torch
from mirrorfusion_b200.config import SD15
from mirrorfusion_b200.schedulers import B200UniPCScheduler
from mirrorfusion_b200.sweep import EvalSweep
from mirrorfusion_b200.synth import make_state_dict
from mirrorfusion_b200.vae import SD_VAE, make_vae_state_dict

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sw = EvalSweep(SD15, make_state_dict(SD15, "unet"), make_state_dict(SD15, "brushnet"), SD_VAE, make_vae_state_dict(SD_VAE, 0, "both"),
               B200UniPCScheduler, H=512, W=512, images_per_call=16, repeats=4, num_inference_steps=steps)
rng = np.random.default_rng(0)
S = 8
rgb = rng.integers(0, 256, (S, 512, 512, 3), dtype=np.uint8)
mask = np.zeros((S, 512, 512), np.uint8); mask[:, 100:350, 120:400] = 255
depth = (rng.random((S, 512, 512), dtype=np.float32) * 4 + 0.5).astype(np.float32)
pe, ne = torch.randn(S, 77, 768), torch.randn(S, 77, 768)
a1, _ = sw.run(rgb, mask, depth, pe, ne, seed=0)
a2, _ = sw.run(rgb, mask, depth, pe, ne, seed=0)
b = np.concatenate([sw.run(rgb, mask, depth, pe, ne, seed=0, rank=r, world=2)[0] for r in range(2)])
d12 = [int(np.abs(a1[i].astype(int) - a2[i].astype(int)).max()) for i in range(len(a1))]
d1b = [int(np.abs(a1[i].astype(int) - b[i].astype(int)).max()) for i in range(len(a1))]
print("run-to-run max |diff| per item:", d12)
print("world 1 vs simulated world 2  :", d1b)
