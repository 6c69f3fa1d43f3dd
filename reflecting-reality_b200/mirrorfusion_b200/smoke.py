"""One small invocation of the hot path on cuda:0, checked against the oracle (driver smoke test)."""
from __future__ import annotations

import torch


def run_smoke(verbose: bool = True) -> float:
    from oracle import mf_oracle as O          # checker only
    from .config import TINY
    from .pipeline import StepEngine
    from .schedulers import B200UniPCScheduler
    from .synth import make_inputs, make_state_dict

    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device; mirrorfusion_b200 has no CPU fallback")
    torch.cuda.set_device(0)
    cfg = TINY
    usd, bsd = make_state_dict(cfg, "unet"), make_state_dict(cfg, "brushnet")
    inp = make_inputs(cfg, 1)
    steps = 3
    eng = StepEngine(cfg, usd, bsd, images=1, H=cfg.sample_size, W=cfg.sample_size, device="cuda:0", use_graph=True)
    eng.set_conditioning(inp["prompt_embeds"].cuda(), inp["conditioning_latents"].cuda())
    got = eng.denoise(inp["latents"].cuda(), B200UniPCScheduler(), steps, 7.5).cpu()
    with torch.no_grad():
        ref = O.denoise_loop(usd, bsd, cfg, O.UniPCOracle(), inp["latents"], inp["prompt_embeds"],
                             inp["conditioning_latents"], steps, 7.5)
    err = ((got - ref).norm() / ref.norm()).item()
    if verbose:
        print(f"[smoke] TINY config, {steps} UniPC steps, CFG 7.5: latents rel-L2 vs fp32 oracle = {err:.3e} "
              f"({eng.launches_per_step} kernel launches/step, CUDA graph)")
    if not (err < 5e-2):
        raise AssertionError(f"smoke parity failed: rel-L2 {err}")
    return err
