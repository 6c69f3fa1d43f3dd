"""LayerNorm throughput on the token shapes of one SD1.5 denoise step at net batch B (CUDA-graph replay, CUDA events)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
from mirrorfusion_b200 import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    ops.lib()
    for T, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
        rows = a.batch * T
        x = torch.randn(rows, C, device="cuda").bfloat16()
        out = torch.empty_like(x)
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        run = lambda: ops.layernorm(x, g, b, out)
        for _ in range(3):
            run()
        side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            run(); side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(a.iters):
                    run()
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"LayerNorm rows={rows:6d} C={C:5d}  {ms * 1e3:7.1f} us  {2 * rows * C * 2 / ms / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
