"""Per-shape throughput of the implicit-GEMM kernel on the conv / linear shapes of one SD1.5 denoise step
 at net batch B.  Prints TFLOP/s per shape; used to steer kernel tuning."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
from mirrorfusion_b200 import ops

bf16 = torch.bfloat16


def time_plan(plan, iters=20):
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        plan.run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--only", type=str, default="", help="substring filter on the shape name")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    ops.lib()
    B = args.batch
    shapes = [  # (name, H, Cin, Cout, ksize)
        ("conv3x3 64x64 320->320", 64, 320, 320, 3), ("conv3x3 64x64 640->320", 64, 640, 320, 3),
        ("conv3x3 64x64 960->320", 64, 960, 320, 3),
        ("conv3x3 32x32 640->640", 32, 640, 640, 3), ("conv3x3 32x32 1280->640", 32, 1280, 640, 3),
        ("conv3x3 32x32 1920->640", 32, 1920, 640, 3),
        ("conv3x3 16x16 1280->1280", 16, 1280, 1280, 3), ("conv3x3 16x16 2560->1280", 16, 2560, 1280, 3),
        ("conv3x3 8x8 1280->1280", 8, 1280, 1280, 3), ("conv3x3 8x8 2560->1280", 8, 2560, 1280, 3),
        ("linear 64x64 320->960 (qkv)", 64, 320, 960, 1), ("linear 64x64 320->2560 (geglu)", 64, 320, 2560, 1),
        ("linear 64x64 1280->320 (ff out)", 64, 1280, 320, 1), ("linear 32x32 640->5120 (geglu)", 32, 640, 5120, 1),
        ("linear 16x16 1280->10240 (geglu)", 16, 1280, 10240, 1), ("zero-conv 64x64 320->320", 64, 320, 320, 1),
        ("linear 64x64 320->320 +res (attn out)", 64, 320, 320, 1), ("linear 32x32 640->640 +res (attn out)", 32, 640, 640, 1),
        ("linear 16x16 1280->1280 +res (attn out)", 16, 1280, 1280, 1), ("linear 64x64 1280->320 +res (ff out)", 64, 1280, 320, 1),
        ("linear 32x32 2560->640 +res (ff out)", 32, 2560, 640, 1), ("linear 32x32 640->1920 (qkv)", 32, 640, 1920, 1),
    ]
    rows = []
    for name, H, cin, cout, ks in shapes:
        if args.only and args.only not in name:
            continue
        x = torch.randn(B, H, H, cin, device="cuda").to(bf16)
        w = (torch.randn(cout, ks * ks * cin, device="cuda") * 0.02).to(bf16)
        geglu = "geglu" in name
        out = torch.empty(B, H, H, cout // 2 if geglu else cout, device="cuda", dtype=bf16)
        bias = torch.zeros(cout, device="cuda")
        res1 = torch.randn(B, H, H, cout, device="cuda").to(bf16) if "+res" in name else None
        plan = ops.ConvPlan(x, w, out, B=B, H=H, W=H, Cin=cin, Cout=cout, ksize=ks, bias=bias, geglu=geglu, res1=res1,
                            block_n=0 if geglu else args.block_n)
        ms = time_plan(plan, args.iters)
        tf = plan.flops / ms / 1e9
        rows.append({"shape": name, "M": B * H * H, "N": cout, "K": ks * ks * cin, "ms": round(ms, 4), "tflops": round(tf, 1)})
        print(f"{name:40s} M={B*H*H:6d} N={cout:5d} K={ks*ks*cin:6d}  {ms:8.4f} ms  {tf:7.1f} TFLOP/s", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if args.only:
        return
    with open(os.path.join(ROOT, "gpurun_out", f"igemm_shapes_b{B}.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
