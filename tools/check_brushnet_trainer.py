"""Per-parameter gradient error of BrushNetTrainer (mirrorfusion_b200/backward.py) against float64 autograd through the oracle's
brushnet_forward, in either precision, plus its forward / backward time (CUDA events).  Diagnostic for the GPU box:

    python tools/check_brushnet_trainer.py --precision bf16 --config tiny     # first bf16 run of the net-level program
    python tools/check_brushnet_trainer.py --precision bf16 --config sd15 --no-check --batch 32 --size 64   # step time at config 4

Writes gpurun_out/brushnet_trainer_<precision>_<config>.json.  The oracle is only the checker."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
import torch
import torch.nn.functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--config", default="tiny", choices=["tiny", "sd15"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=16)
    ap.add_argument("--no-check", action="store_true", help="skip the float64 autograd comparison (large configs): timing only")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    from mirrorfusion_b200 import ops
    from mirrorfusion_b200.backward import BrushNetTrainer, brushnet_resnet_prefixes, brushnet_shapes, pack_brushnet, unpack_conv_grad
    from mirrorfusion_b200.config import SD15, TINY
    from mirrorfusion_b200.synth import make_state_dict
    from mirrorfusion_b200.train import FlatParams
    ops.lib()
    cfg = TINY if args.config == "tiny" else SD15
    B, H, W = args.batch, args.size, args.size
    gen = torch.Generator().manual_seed(10)
    sd32 = make_state_dict(cfg, "brushnet")
    sample = torch.randn(B, cfg.in_channels, H, W, generator=gen)
    cond = torch.randn(B, cfg.conditioning_channels, H, W, generator=gen)
    t = torch.randint(0, 1000, (B,), generator=gen)
    shapes = brushnet_shapes(cfg)
    flat = FlatParams(shapes, "cuda")
    flat.load_state_dict(pack_brushnet(cfg, sd32))
    net = BrushNetTrainer(flat, cfg, B=B, H=H, W=W, precision=args.precision)
    dt = torch.float32 if args.precision == "fp32" else torch.bfloat16
    td, tm, tu = net.forward(sample.cuda(), cond.cuda(), t.cuda())
    taps = td + [tm] + tu
    d_taps = [torch.randn(x.shape, generator=gen).to(dt).cuda() for x in taps]
    res = {"precision": args.precision, "config": args.config, "B": B, "size": H, "params": int(sum(v.numel() for v in sd32.values()))}

    def run_bwd():
        flat.grad.zero_()
        net.backward(d_taps[:len(td)], d_taps[len(td)], d_taps[len(td) + 1:])

    run_bwd()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(args.iters):
        net.forward(sample.cuda(), cond.cuda(), t.cuda())
    ev[1].record()
    for _ in range(args.iters):
        run_bwd()
    ev[2].record()
    torch.cuda.synchronize()
    res["forward_ms"] = ev[0].elapsed_time(ev[1]) / args.iters
    res["backward_ms"] = ev[1].elapsed_time(ev[2]) / args.iters
    res["max_memory_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    print(f"forward {res['forward_ms']:.2f} ms, backward {res['backward_ms']:.2f} ms (eager launches, B={B}, {H}x{W}); "
          f"peak device memory {res['max_memory_gb']:.1f} GiB")

    if not args.no_check:
        from oracle import mf_oracle as O
        sd = {k: v.double().requires_grad_(True) for k, v in sd32.items()}
        down, mid, up = O.brushnet_forward(sd, cfg, sample.double(), t, cond.double())
        ref_taps = down + [mid] + up
        nchw = lambda g, r: g.float().cpu().reshape(B, r.shape[2], r.shape[3], r.shape[1]).permute(0, 3, 1, 2).double()
        sum((a * nchw(g, a)).sum() for a, g in zip(ref_taps, d_taps)).backward()
        rel = lambda a, b: float(a.double().cpu().norm()) if float(b.norm()) < 1e-9 else float((a.double().cpu() - b.double()).norm() / b.double().norm())
        res["tap_err_max"] = max(rel(nchw(a, r), r.detach()) for a, r in zip(taps, ref_taps))
        prefixes = brushnet_resnet_prefixes(cfg)
        errs = {}
        for name in shapes:
            if name.endswith(".weight.b"):
                continue
            if name == "time_emb_proj.wcat":
                want, got = torch.cat([sd[p + ".time_emb_proj.weight"].grad for p in prefixes], 0), flat.g(name)
            elif name == "time_emb_proj.bcat":
                want, got = torch.cat([sd[p + ".time_emb_proj.bias"].grad for p in prefixes], 0), flat.g(name)
            elif name.endswith(".weight.a"):
                want, got = sd[name[:-2]].grad[:, :, 0, 0], torch.cat([flat.g(name), flat.g(name[:-2] + ".b")], 1)
            else:
                want, got = sd[name].grad, flat.g(name)
                got = unpack_conv_grad(got, 3) if want.dim() == 4 and want.shape[-1] == 3 else got.reshape(want.shape)
            errs[name] = rel(got, want)
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
        res["grad_err_max"], res["grad_err_median"] = worst[0][1], sorted(errs.values())[len(errs) // 2]
        res["worst"] = worst
        print(f"taps max rel-L2 {res['tap_err_max']:.3e}; parameter gradients: max {res['grad_err_max']:.3e}, median {res['grad_err_median']:.3e}")
        for k, v in worst:
            print(f"  {v:.3e}  {k}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"brushnet_trainer_{args.precision}_{args.config}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
