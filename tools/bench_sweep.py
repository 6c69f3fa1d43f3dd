"""(Launch with torchrun for N > 1: ranks shard the item list; rank 0 prints one JSON line with the aggregate rate — max time
over ranks — and a SHA-256 of all images in global item order, which must not depend on N: BASELINE config 3.)
End-to-end images/s of the batched eval sweep (sweep.EvalSweep) at the real geometry: SD1.5 nets + SD VAE, 512x512,
50 UniPC steps, CFG 7.5, uint8 host arrays in -> uint8 host arrays out (preprocessing, VAE encode, denoise loop, VAE decode,
postprocessing all inside the timed region).  Synthetic SynMirror-shaped inputs, random-init weights.  Prints one JSON line."""
import argparse
import hashlib
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
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sw = EvalSweep(SD15, make_state_dict(SD15, "unet"), make_state_dict(SD15, "brushnet"), SD_VAE, make_vae_state_dict(SD_VAE, 0, "both"),
                   B200UniPCScheduler, H=512, W=512, images_per_call=a.images_per_call, repeats=a.repeats, num_inference_steps=a.steps)
    rng = np.random.default_rng(0)
    S = a.samples
    rgb = rng.integers(0, 256, (S, 512, 512, 3), dtype=np.uint8)
    mask = np.zeros((S, 512, 512), np.uint8)
    mask[:, 100:350, 120:400] = 255
    depth = (rng.random((S, 512, 512), dtype=np.float32) * 4 + 0.5).astype(np.float32)
    g = torch.Generator().manual_seed(99)          # every rank must build the SAME sweep inputs
    pe, ne = torch.randn(S, 77, 768, generator=g), torch.randn(S, 77, 768, generator=g)
    sw.run(rgb[:4], mask[:4], depth[:4], pe[:4], ne[:4], seed=1)           # warm-up: graph capture, lazy kernel attributes
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out, items = sw.run(rgb, mask, depth, pe, ne, seed=0, rank=rank, world=world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_items = S * a.repeats
    if world > 1:
        tmax = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = tmax.item()
        from mirrorfusion_b200.sharding import shard_range
        counts = [len(shard_range(n_items, r, world)) for r in range(world)]
        pad = torch.zeros(max(counts), 512, 512, 3, dtype=torch.uint8, device="cuda")
        pad[: out.shape[0]] = torch.from_numpy(out).cuda()
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)                                       # the path's only collective: the final gather
        out = torch.cat([p_[:c] for p_, c in zip(parts, counts)]).cpu().numpy()
    if rank == 0:
        print(json.dumps({"metric": "eval_sweep_images_per_s_512x512_50_unipc_steps_cfg7.5_incl_vae_and_io", "value": n_items / dt,
                          "unit": "images/s", "n_gpus": world, "images": n_items, "seconds": dt, "images_per_call": a.images_per_call,
                          "steps": a.steps, "out_shape": list(out.shape), "sha256_of_all_images": hashlib.sha256(out.tobytes()).hexdigest()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
