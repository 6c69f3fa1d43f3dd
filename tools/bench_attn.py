"""Per-shape throughput of the flash-attention kernel on the attention shapes of one SD1.5 denoise step
 at net batch B.  Prints ms and TFLOP/s (unpadded head dim) per shape; used to steer tuning."""
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
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    ops.lib()
    B = args.batch
    shapes = [  # (name, Tq, Tk, heads, head_dim)
        ("self 64x64 d40", 4096, 4096, 8, 40), ("cross 64x64 d40", 4096, 77, 8, 40),
        ("self 32x32 d80", 1024, 1024, 8, 80), ("cross 32x32 d80", 1024, 77, 8, 80),
        ("self 16x16 d160", 256, 256, 8, 160), ("cross 16x16 d160", 256, 77, 8, 160),
        ("self 8x8 d160", 64, 64, 8, 160),
    ]
    for name, Tq, Tk, H, D in shapes:
        if args.only and args.only not in name:
            continue
        C = H * D
        q = torch.randn(B, Tq, C, device="cuda").to(bf16)
        k = torch.randn(B, Tk, C, device="cuda").to(bf16)
        v = torch.randn(B, Tk, C, device="cuda").to(bf16)
        out = torch.empty(B, Tq, C, device="cuda", dtype=bf16)
        run = lambda: ops.attention(q, k, v, out, B=B, heads=H, head_dim=D, Tq=Tq, Tk=Tk)
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
        fl = 4.0 * B * H * Tq * Tk * D
        print(f"{name:20s} Tq={Tq:5d} Tk={Tk:5d} H={H} D={D:3d}  {ms:8.4f} ms  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
