"""Throughput of the fine-tune step's glue kernels (csrc/train.cu) at BASELINE config 4's sizes: batch 32 per GPU, latents
4x64x64, 618.8 M trainable BrushNet parameters in flat buffers.  CUDA events around `iters` back-to-back launches after warm-up
(the buffers are far larger than L2 for the optimizer; the small kernels are reported as launch-latency numbers).
Algorithmic bytes per element: AdamW 16 read (p, g, m, v) + 14 written (p, m, v, bf16 copy) = 30; sqnorm 4; add_noise 12; MSE 12.
Writes gpurun_out/train_glue_bench.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
from mirrorfusion_b200 import ops
from mirrorfusion_b200.train import B200AdamW, FlatParams, NoiseSchedule, TrainLoss

bf16 = torch.bfloat16


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--params", type=int, default=618_832_960)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--skip-wgrad", action="store_true")
    ap.add_argument("--only-conv", action="store_true", help="only the conv forward / dgrad / wgrad section (ncu captures)")
    args = ap.parse_args()
    ops.lib()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    res = {"params": args.params, "batch": args.batch, "hbm_peak_gbs": peaks["hbm_gbs"], "kernels": {}}

    def rec(name, ms, nbytes=None, flops=None, note=""):
        r = {"ms": round(ms, 4)}
        if nbytes:
            r["gbs"] = round(nbytes / ms / 1e6, 1)
            r["frac_hbm"] = round(r["gbs"] / peaks["hbm_gbs"], 3)
        if flops:
            r["tflops"] = round(flops / ms / 1e9, 1)
        if note:
            r["note"] = note
        res["kernels"][name] = r
        print(name, r, flush=True)

    n = args.params if not args.only_conv else 1024
    flat = FlatParams({"all": (n,)}, "cuda")
    flat.param.normal_(0, 0.02)
    flat.grad.normal_(0, 1e-3)
    opt = B200AdamW(flat, lr=5e-6)
    rec("clip+adamw+bf16 (1 launch + sqnorm)", timed(lambda: opt.step(max_grad_norm=1.0), args.iters), nbytes=34.0 * n,
        note="34 B/elem: 4 (norm pass) + 30 (update)")
    rec("adamw+bf16 only", timed(lambda: opt.step(), args.iters), nbytes=30.0 * n)
    ws, sq = torch.zeros(ops.SQNORM_WS_FLOATS, device="cuda"), torch.zeros(1, device="cuda")
    rec("grad_sqnorm", timed(lambda: ops.grad_sqnorm(flat.grad, ws, sq), args.iters), nbytes=4.0 * n)
    del flat, opt
    torch.cuda.empty_cache()

    B, m = args.batch, 4 * 64 * 64
    ns = NoiseSchedule("cuda")
    x0, noise, out = torch.randn(B, m, device="cuda"), torch.randn(B, m, device="cuda"), torch.empty(B, m, device="cuda")
    t = ns.sample_timesteps(B, torch.Generator().manual_seed(0)).cuda()
    rec("add_noise", timed(lambda: ns.add_noise(x0, noise, t, out=out), 50), nbytes=12.0 * B * m, note="2 MB per tensor: launch latency")
    L = TrainLoss(B, "cuda")
    grad = torch.empty_like(x0)
    rec("mse_loss+grad (2 launches)", timed(lambda: L(x0, noise, grad=grad), 50), nbytes=12.0 * B * m, note="launch latency")

    # one resnet conv of the 64x64 level at batch 32: forward, data gradient (same tcgen05 kernel), weight gradient (CUDA cores)
    H = W = 64
    Cin = Cout = 320
    x = torch.randn(B, H, W, Cin, device="cuda").to(bf16)
    dy = torch.randn(B, H, W, Cout, device="cuda").to(bf16)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (9 * Cin) ** -0.5
    y, dx = torch.empty_like(dy), torch.empty_like(x)
    fwd = ops.ConvPlan(x, ops.pack_conv_weight(w), y, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3)
    bwd = ops.ConvPlan(dy, ops.pack_conv_dgrad_weight(w), dx, B=B, H=H, W=W, Cin=Cout, Cout=Cin, ksize=3)
    fl = 2.0 * B * H * W * Cout * 9 * Cin
    rec("conv3x3 320->320 @64x64 forward (tcgen05)", timed(fwd.run, 20), flops=fl)
    rec("conv3x3 320->320 @64x64 dgrad (tcgen05, same kernel)", timed(bwd.run, 20), flops=fl)
    dw, db = torch.empty(Cout, 9 * Cin, device="cuda"), torch.empty(Cout, device="cuda")
    wws = torch.empty(ops.conv_wgrad_ws_floats(B, H, W, Cin, Cout, 3), device="cuda")
    rec("conv3x3 320->320 @64x64 wgrad + dbias (tensor cores, split-K mma.sync; 4 launches)",
        timed(lambda: ops.conv_wgrad(x, dy, dw, db, B=B, H=H, W=W, ksize=3, ws=wws), 20), flops=fl)
    rec("conv3x3 320->320 @64x64 wgrad only (tensor cores; 2 launches)",
        timed(lambda: ops.conv_wgrad(x, dy, dw, None, B=B, H=H, W=W, ksize=3, ws=wws), 20), flops=fl)
    if not args.skip_wgrad:
        rec("conv3x3 320->320 @64x64 wgrad (CUDA cores, first version)",
            timed(lambda: ops.conv_wgrad(x, dy, dw, db, B=B, H=H, W=W, ksize=3, cuda_cores=True), 3, warm=1), flops=fl)
    gamma, beta = torch.ones(Cin, device="cuda"), torch.zeros(Cin, device="cuda")
    gws = torch.zeros(2 * B * Cin, device="cuda")
    dg, dbt = torch.empty(Cin, device="cuda"), torch.empty(Cin, device="cuda")
    rec("groupnorm+silu backward 320ch @64x64 (first version)",
        timed(lambda: ops.groupnorm_bwd(x, None, dy, gamma, beta, dx, None, gws, B=B, HW=H * W, groups=32, eps=1e-5, silu=True,
                                        dgamma=dg, dbeta=dbt), 5, warm=1), nbytes=6.0 * B * H * W * Cin,
        note="algorithmic 6 B/elem (x, dy read, dx written, bf16)")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "train_glue_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
