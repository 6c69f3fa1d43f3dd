"""Step time of the fp32 PARITY MODE at BASELINE config 1's geometry (SD1.5 nets, 1 image + CFG, 64x64 latents).  The mode
is a correctness instrument (CUDA-core FFMA kernels); this number only says how long a parity run takes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
from mirrorfusion_b200.config import SD15
from mirrorfusion_b200.pipeline import StepEngine
from mirrorfusion_b200.schedulers import B200DDIMScheduler
from mirrorfusion_b200.synth import make_inputs, make_state_dict

usd, bsd = make_state_dict(SD15, "unet"), make_state_dict(SD15, "brushnet")
inp = make_inputs(SD15, 1)
for prec in ("fp32", "bf16"):
    eng = StepEngine(SD15, usd, bsd, 1, 64, 64, precision=prec)
    eng.set_conditioning(inp["prompt_embeds"].cuda(), inp["conditioning_latents"].cuda())
    eng.denoise(inp["latents"].cuda(), B200DDIMScheduler(), 4, 7.5)          # warm-up + graph capture
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.denoise(inp["latents"].cuda(), B200DDIMScheduler(), 4, 7.5)
    torch.cuda.synchronize()
    print(f"{prec}: {(time.perf_counter() - t0) / 4 * 1e3:8.2f} ms per step (config 1: 1 image, 4 DDIM steps, CFG 7.5)", flush=True)
    del eng
