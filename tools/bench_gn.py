"""Per-shape throughput of GroupNorm(+SiLU) (statistics pass + apply pass) on the shapes of one SD1.5 denoise step
 at net batch B: ms and effective GB/s (algorithmic traffic = 2 reads + 1 write of the tensor)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
from mirrorfusion_b200 import ops

bf16 = torch.bfloat16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--only", type=str, default="")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    ops.lib()
    B = args.batch
    shapes = [("64x64 320", 64, 320, 0), ("64x64 640+320", 64, 640, 320), ("64x64 320+320", 64, 320, 320),
              ("32x32 640", 32, 640, 0), ("32x32 1280+640", 32, 1280, 640), ("16x16 1280", 16, 1280, 0),
              ("16x16 1280+1280", 16, 1280, 1280), ("8x8 1280", 8, 1280, 0), ("8x8 1280+1280", 8, 1280, 1280)]
    for name, H, C1, C2 in shapes:
        if args.only and args.only not in name:
            continue
        HW, C = H * H, C1 + C2
        x1 = torch.randn(B, HW, C1, device="cuda").to(bf16)
        x2 = torch.randn(B, HW, C2, device="cuda").to(bf16) if C2 else None
        out = torch.empty(B, HW, C, device="cuda", dtype=bf16)
        gamma = torch.ones(C, device="cuda")
        beta = torch.zeros(C, device="cuda")
        ws = torch.zeros(ops.gn_ws_floats(B, 32), device="cuda")
        run = lambda: ops.groupnorm(x1, x2, gamma, beta, out, ws, B=B, HW=HW, groups=32, eps=1e-5, silu=True)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        # replay from a CUDA graph: the python/ctypes launch path (~10 us) would otherwise hide the small shapes
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            run()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(args.iters):
                    run()
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        nbytes = 3.0 * B * HW * C * 2
        print(f"GN+SiLU {name:18s} {B * HW * C * 2 / 1e6:7.1f} MB  {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
