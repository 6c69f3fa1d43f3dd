"""End-to-end images/s of the batched eval sweep (sweep.EvalSweep) at the real geometry: SD1.5 nets + SD VAE, 512x512,
50 UniPC steps, CFG 7.5, uint8 host arrays in -> uint8 host arrays out (preprocessing, VAE encode, denoise loop, VAE decode,
postprocessing all inside the timed region).  Synthetic SynMirror-shaped inputs, random-init weights.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import numpy as np
import torch
from mirrorfusion_b200.config import SD15
from mirrorfusion_b200.schedulers import B200UniPCScheduler
from mirrorfusion_b200.sweep import EvalSweep
from mirrorfusion_b200.synth import make_state_dict
from mirrorfusion_b200.vae import SD_VAE, make_vae_state_dict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images-per-call", type=int, default=16)
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--repeats", type=int, default=4)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    sw = EvalSweep(SD15, make_state_dict(SD15, "unet"), make_state_dict(SD15, "brushnet"), SD_VAE, make_vae_state_dict(SD_VAE, 0, "both"),
                   B200UniPCScheduler, H=512, W=512, images_per_call=a.images_per_call, repeats=a.repeats, num_inference_steps=a.steps)
    rng = np.random.default_rng(0)
    S = a.samples
    rgb = rng.integers(0, 256, (S, 512, 512, 3), dtype=np.uint8)
    mask = np.zeros((S, 512, 512), np.uint8)
    mask[:, 100:350, 120:400] = 255
    depth = (rng.random((S, 512, 512), dtype=np.float32) * 4 + 0.5).astype(np.float32)
    pe, ne = torch.randn(S, 77, 768), torch.randn(S, 77, 768)
    sw.run(rgb[:4], mask[:4], depth[:4], pe[:4], ne[:4], seed=1)           # warm-up: graph capture, lazy kernel attributes
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, items = sw.run(rgb, mask, depth, pe, ne, seed=0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"metric": "eval_sweep_images_per_s_512x512_50_unipc_steps_cfg7.5_incl_vae_and_io", "value": len(items) / dt,
                      "unit": "images/s", "images": len(items), "seconds": dt, "images_per_call": a.images_per_call,
                      "steps": a.steps, "out_shape": list(out.shape), "out_mean": float(out.mean())}))


if __name__ == "__main__":
    main()
