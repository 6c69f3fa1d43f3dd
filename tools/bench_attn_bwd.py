"""Time the tcgen05 attention forward (with LSE) and backward kernels on the self- / cross-attention shapes of the SD1.5 UNet:
python tools/bench_attn_bwd.py [--batch 8] [--only "64x64"]   (ms per call, TFLOP/s: 4 T S C forward, 2.5x that backward)."""
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
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    ops.lib()
    B, heads = a.batch, 8
    shapes = [("self 64x64", 4096, 4096, 40), ("self 32x32", 1024, 1024, 80), ("self 16x16", 256, 256, 160),
              ("cross 64x64", 4096, 77, 40), ("cross 32x32", 1024, 77, 80)]
    for name, Tq, Tk, d in shapes:
        if a.only and a.only not in name:
            continue
        C = heads * d
        cross = Tk != Tq
        if cross:
            q = torch.randn(B, Tq, C, device="cuda").to(bf16)
            k = torch.randn(B, Tk, C, device="cuda").to(bf16)
            v = torch.randn(B, Tk, C, device="cuda").to(bf16)
            ld = dict(ldq=C, ldk=C, ldv=C)
            dq, dk, dv, ldd = torch.empty_like(q), None, None, dict(lddq=C)
        else:
            qkv = torch.randn(B, Tq, 3 * C, device="cuda").to(bf16)
            q, k, v = qkv.view(-1), qkv.view(-1)[C:], qkv.view(-1)[2 * C:]
            ld = dict(ldq=3 * C, ldk=3 * C, ldv=3 * C)
            dqkv = torch.empty_like(qkv)
            dq, dk, dv = dqkv.view(-1), dqkv.view(-1)[C:], dqkv.view(-1)[2 * C:]
            ldd = dict(lddq=3 * C, lddk=3 * C, lddv=3 * C)
        o = torch.empty(B, Tq, C, device="cuda", dtype=bf16)
        do = torch.randn(B, Tq, C, device="cuda").to(bf16)
        lse = torch.zeros(B * heads * Tq, device="cuda")
        dvec = torch.zeros_like(lse)
        geo = dict(B=B, heads=heads, head_dim=d, Tq=Tq, Tk=Tk)
        fwd = lambda: ops.attention_lse(q, k, v, o, lse, ldo=C, **ld, **geo)
        bwd = lambda: ops.attention_bwd(q, k, v, o, do, lse, dvec, dq, dk, dv, ldo=C, lddo=C, **ld, **ldd, **geo)
        res = []
        for fn in (fwd, bwd):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / a.iters)
        fl = 4.0 * B * Tq * Tk * C
        print(f"{name:12s} B={B}: forward {res[0]:.3f} ms ({fl / res[0] / 1e9:.0f} TFLOP/s)   backward {res[1]:.3f} ms "
              f"({(1.5 if cross else 2.5) * fl / res[1] / 1e9:.0f} TFLOP/s)   bwd/fwd {res[1] / res[0]:.2f}")


if __name__ == "__main__":
    main()
