"""Per-launch breakdown of one denoise step (CUDA events around every program entry, eager): igemm launches grouped by
shape with their time, algorithmic TFLOP/s and the time they would take at the measured sustained peak — shows where
the step's remaining tensor-core headroom sits."""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
from mirrorfusion_b200.config import SD15
from mirrorfusion_b200.pipeline import StepEngine
from mirrorfusion_b200.schedulers import B200UniPCScheduler
from mirrorfusion_b200.synth import make_inputs, make_state_dict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--peak", type=float, default=1417.3, help="TFLOP/s used for the 'ideal' column")
    args = ap.parse_args()
    cfg, H = SD15, args.latent
    usd, bsd = make_state_dict(cfg, "unet"), make_state_dict(cfg, "brushnet")
    inp = make_inputs(cfg, args.images, height=H, width=H)
    eng = StepEngine(cfg, usd, bsd, args.images, H, H, "cuda", use_graph=False)
    eng.set_conditioning(inp["prompt_embeds"].cuda(), inp["conditioning_latents"].cuda())
    eng.x.copy_(inp["latents"].cuda())
    sched = B200UniPCScheduler()
    sched.set_timesteps(50)
    table = sched.coefficient_table(7.5).cuda()
    for i in range(2):
        eng.step(float(sched.timesteps[i]), table[i], 1.0)
    torch.cuda.synchronize()
    rows = []
    for _ in range(3):                          # keep the last of 3 passes (warm)
        rows = []
        for net, e in (("brushnet", eng.bn), ("unet", eng.unet)):
            rows += [(net,) + r for r in e.run_timed(per_entry=True, skip=e.n_time_ops)]   # timestep path: hoisted in the real loop
    agg = collections.OrderedDict()
    fam = collections.Counter()
    for net, tag, note, ms, fl in rows:
        fam[tag] += ms
        if tag != "igemm":
            continue
        a = agg.setdefault(note, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += fl
    print("families (ms, eager events incl. launch gaps):", {k: round(v, 3) for k, v in fam.items()})
    for tag in ("misc", "attention", "groupnorm", "layernorm"):
        ent = sorted([(ms, net, i) for i, (net, t, note, ms, fl) in enumerate(rows) if t == tag], reverse=True)
        print(f"{tag}: {len(ent)} entries; top:", [(round(ms * 1e3, 1), net, i) for ms, net, i in ent[:12]], "(us, net, program index)")
    tot = sum(a[1] for a in agg.values())
    print(f"igemm: {tot:.3f} ms in {sum(a[0] for a in agg.values())} launches; by shape, sorted by headroom (ms above the ideal):")
    out = []
    for note, (n, ms, fl) in agg.items():
        ideal = fl / args.peak / 1e9
        out.append((ms - ideal, note, n, ms, fl / ms / 1e9 if ms > 0 else 0.0, ideal))
    for head, note, n, ms, tf, ideal in sorted(out, reverse=True):
        print(f"  {note:44s} x{n:3d}  {ms:7.3f} ms  {tf:7.1f} TFLOP/s  ideal {ideal:6.3f}  headroom {head:6.3f}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "profile_step.json"), "w") as f:
        json.dump([{"shape": o[1], "launches": o[2], "ms": o[3], "tflops": o[4], "ideal_ms": o[5]} for o in sorted(out, reverse=True)], f, indent=1)


if __name__ == "__main__":
    main()
