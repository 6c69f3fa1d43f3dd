"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only mount)

The reference (the diffusers 0.27 fork under /root/reference/MirrorFusion/src) is imported
read-only; its `UNet2DConditionModel`, `BrushNetModel`, `DDIMScheduler` and
`UniPCMultistepScheduler` are instantiated, loaded (strict) with the seeded synthetic
state_dicts from `mirrorfusion_b200.synth`, and run on CPU in fp32.  The vectors it writes
are what `tests/test_oracle_golden.py` pins `oracle/mf_oracle.py` against, and what the
GPU parity tests compare the CUDA path with (the GPU box has no /root/reference).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reflecting-reality_b200"))
REF_SRC = "/root/reference/MirrorFusion/src"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    import transformers.utils as tu  # shim: symbol removed in transformers 5, imported by pipeline_loading_utils.py:44
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    sys.path.insert(0, REF_SRC)
    import diffusers  # noqa
    return diffusers


def build_reference_nets(diffusers, cfg, seed=0):
    from mirrorfusion_b200.synth import make_state_dict
    n = len(cfg.block_out_channels)
    down = tuple("CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg.down_has_attn)
    up = tuple("CrossAttnUpBlock2D" if a else "UpBlock2D" for a in cfg.up_has_attn)
    unet = diffusers.UNet2DConditionModel(
        sample_size=cfg.sample_size, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
        block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
        down_block_types=down, up_block_types=up, cross_attention_dim=cfg.cross_attention_dim,
        attention_head_dim=cfg.heads, norm_num_groups=cfg.norm_num_groups).eval()
    bn = diffusers.BrushNetModel.from_unet(unet, conditioning_channels=cfg.conditioning_channels).eval()
    # from_unet ALIASES the bias Parameter (`brushnet.conv_in_condition.bias = unet.conv_in.bias`,
    # S/models/brushnet.py:518): loading one net's state_dict would silently overwrite the other's.
    # Un-share it so both nets carry exactly the seeded tensors.
    bn.conv_in_condition.bias = torch.nn.Parameter(bn.conv_in_condition.bias.detach().clone())
    usd = make_state_dict(cfg, "unet", seed)
    bsd = make_state_dict(cfg, "brushnet", seed)
    unet.load_state_dict(usd, strict=True)
    bn.load_state_dict(bsd, strict=True)
    assert torch.equal(unet.conv_in.bias.detach(), usd["conv_in.bias"])
    return unet, bn, usd, bsd


@torch.no_grad()
def ref_step(unet, bn, x, t, ehs, cond, scale=1.0):
    d, m, u = bn(x, t, encoder_hidden_states=ehs, brushnet_cond=cond, conditioning_scale=scale, return_dict=False)
    eps = unet(x, t, encoder_hidden_states=ehs, down_block_add_samples=[a.clone() for a in d],
               mid_block_add_sample=m, up_block_add_samples=[a.clone() for a in u], return_dict=False)[0]
    return eps, d, m, u


def golden_step(diffusers, cfg, name, images=1, t=500, full_taps=True, seed=0, scale=1.0):
    from mirrorfusion_b200.synth import make_inputs
    unet, bn, _, _ = build_reference_nets(diffusers, cfg, seed)
    inp = make_inputs(cfg, images)
    x = torch.cat([inp["latents"]] * 2)
    t0 = time.time()
    eps, d, m, u = ref_step(unet, bn, x, torch.tensor(t), inp["prompt_embeds"], inp["conditioning_latents"], scale)
    dt = time.time() - t0
    plain = unet(x, torch.tensor(t), encoder_hidden_states=inp["prompt_embeds"], return_dict=False)[0]
    out = {"noise_pred": eps.numpy(), "noise_pred_no_taps": plain.detach().numpy(), "t": np.int64(t),
           "images": np.int64(images), "seed": np.int64(seed), "scale": np.float64(scale)}
    taps = list(d) + [m] + list(u)
    if full_taps:
        for k, a in enumerate(taps):
            out[f"tap{k:02d}"] = a.numpy()
    else:  # full-size config: keep the fixtures small -> per-tap statistics + a strided sample
        for k, a in enumerate(taps):
            out[f"tap{k:02d}_l2"] = np.float64(a.double().norm().item())
            out[f"tap{k:02d}_sample"] = a.flatten()[:: max(1, a.numel() // 4096)][:4096].numpy()
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(f"{name}: ref step {dt:.2f}s  |eps|={eps.norm():.4f} rel(taps effect)={(eps - plain).norm() / plain.norm():.3f}")


def golden_loop(diffusers, cfg, name, sched_kind, steps, images=1, seed=0):
    from mirrorfusion_b200.synth import make_inputs
    unet, bn, _, _ = build_reference_nets(diffusers, cfg, seed)
    inp = make_inputs(cfg, images)
    base = diffusers.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                   clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    sched = base if sched_kind == "ddim" else diffusers.UniPCMultistepScheduler.from_config(base.config)
    sched.set_timesteps(steps)
    lat = inp["latents"] * sched.init_noise_sigma
    lats, epss = [], []
    for t in sched.timesteps:                                      # pipeline_brushnet.py:1250-1315
        x = sched.scale_model_input(torch.cat([lat] * 2), t)
        eps, *_ = ref_step(unet, bn, x, t, inp["prompt_embeds"], inp["conditioning_latents"])
        u_, c_ = eps.chunk(2)
        guided = u_ + 7.5 * (c_ - u_)
        lat = sched.step(guided, t, lat, return_dict=False)[0]
        lats.append(lat.numpy().copy())
        epss.append(eps.numpy().copy())
    np.savez_compressed(os.path.join(GOLD, name), latents=np.stack(lats), noise_pred=np.stack(epss),
                        timesteps=sched.timesteps.numpy(), steps=np.int64(steps), guidance=np.float64(7.5))
    print(f"{name}: {steps} {sched_kind} steps, final |lat|={lat.norm():.4f}")


def golden_sched_only(diffusers, name):
    """Scheduler-only known-answer loops on a fixed pseudo-model (in the spirit of
    T/schedulers/test_schedulers.py:326-368 dummy_model / dummy_sample_deter)."""
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(2, 4, 8, 8, generator=g)
    out = {"x0": x0.numpy()}
    base = diffusers.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                   clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    for kind, n in (("ddim", 4), ("ddim", 10), ("unipc", 5), ("unipc", 10), ("unipc", 50)):
        s = base if kind == "ddim" else diffusers.UniPCMultistepScheduler.from_config(base.config)
        s.set_timesteps(n)
        x = x0.clone()
        traj = []
        for i, t in enumerate(s.timesteps):
            eps = torch.sin(3.0 * x + 0.01 * float(t)) * 0.9 + 0.1 * x      # deterministic pseudo-model
            x = s.step(eps, t, x, return_dict=False)[0]
            traj.append(x.numpy().copy())
        out[f"{kind}{n}_traj"] = np.stack(traj)
        out[f"{kind}{n}_timesteps"] = s.timesteps.numpy()
        if kind == "unipc":
            out[f"{kind}{n}_sigmas"] = s.sigmas.numpy()
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(f"{name}: scheduler trajectories written")


def golden_step_guess_mode(diffusers, cfg, name, images=2, t=500, seed=0, scale=0.9):
    """guess_mode with CFG as the reference pipeline runs it (pipeline_brushnet.py:1262-1301): BrushNetModel.forward(guess_mode=True)
    on the conditional half, zeros concatenated for the unconditional half, then the UNet."""
    from mirrorfusion_b200.synth import make_inputs
    unet, bn, _, _ = build_reference_nets(diffusers, cfg, seed)
    inp = make_inputs(cfg, images)
    x = torch.cat([inp["latents"]] * 2)
    n = images
    with torch.no_grad():
        d, m, u = bn(x[n:], torch.tensor(t), encoder_hidden_states=inp["prompt_embeds"][n:], brushnet_cond=inp["conditioning_latents"][n:],
                     conditioning_scale=scale, guess_mode=True, return_dict=False)
        d = [torch.cat([torch.zeros_like(a), a]) for a in d]
        m = torch.cat([torch.zeros_like(m), m])
        u = [torch.cat([torch.zeros_like(a), a]) for a in u]
        eps = unet(x, torch.tensor(t), encoder_hidden_states=inp["prompt_embeds"], down_block_add_samples=[a.clone() for a in d],
                   mid_block_add_sample=m, up_block_add_samples=[a.clone() for a in u], return_dict=False)[0]
    out = {"noise_pred": eps.numpy(), "t": np.int64(t), "images": np.int64(images), "seed": np.int64(seed), "scale": np.float64(scale)}
    for k, a in enumerate(list(d) + [m] + list(u)):
        out[f"tap{k:02d}_l2"] = np.float64(a.double().norm().item())
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(f"{name}: guess-mode step |eps|={eps.norm():.4f}")


def golden_sched_eta(diffusers, name):
    """Stochastic DDIM (eta > 0): the reference's DDIMScheduler.step draws its variance noise from `generator` with randn_tensor
    (scheduling_ddim.py:452-464); trajectory on the pseudo-model of golden_sched_only."""
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(2, 4, 8, 8, generator=g)
    out = {"x0": x0.numpy()}
    for eta, n, seed in ((0.7, 6, 11), (1.0, 10, 5)):
        s = diffusers.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                                    set_alpha_to_one=False, steps_offset=1)
        s.set_timesteps(n)
        gen = torch.Generator().manual_seed(seed)
        x = x0.clone()
        traj = []
        for t in s.timesteps:
            eps = torch.sin(3.0 * x + 0.01 * float(t)) * 0.9 + 0.1 * x
            x = s.step(eps, t, x, eta=eta, generator=gen, return_dict=False)[0]
            traj.append(x.numpy().copy())
        out[f"eta{eta}_n{n}_seed{seed}_traj"] = np.stack(traj)
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(f"{name}: stochastic DDIM trajectories written")


def golden_psnr(diffusers, steps=20):
    """Final-image protocol of north_star: run the REFERENCE loop (SD1.5-shaped random-init nets, UniPC, CFG 7.5, fp32 CPU),
    keep the final latents, and trace a small random-init reference AutoencoderKL decoder to TorchScript so that the
    GPU box (no reference there) can decode both sides with the same VAE and measure PSNR on uint8 images."""
    from mirrorfusion_b200.config import SD15
    from mirrorfusion_b200.synth import make_inputs
    torch.manual_seed(123)
    vae = diffusers.AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 2,
                                  up_block_types=("UpDecoderBlock2D",) * 2, block_out_channels=(32, 64), layers_per_block=1,
                                  latent_channels=4, norm_num_groups=8, sample_size=128, scaling_factor=0.18215).eval()

    class Dec(torch.nn.Module):
        def __init__(self, v):
            super().__init__()
            self.v = v

        def forward(self, z):
            return self.v.decode(z / 0.18215, return_dict=False)[0]          # pipeline_brushnet.py:1342

    with torch.no_grad():
        ts = torch.jit.trace(Dec(vae), torch.randn(1, 4, 64, 64), check_trace=False)
    ts.save(os.path.join(GOLD, "tiny_vae_decoder.pt"))
    unet, bn, _, _ = build_reference_nets(diffusers, SD15, 0)
    inp = make_inputs(SD15, 1)
    base = diffusers.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                   clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    sched = diffusers.UniPCMultistepScheduler.from_config(base.config)
    sched.set_timesteps(steps)
    lat = inp["latents"] * sched.init_noise_sigma
    t0 = time.time()
    for t in sched.timesteps:
        x = torch.cat([lat] * 2)
        eps, *_ = ref_step(unet, bn, x, t, inp["prompt_embeds"], inp["conditioning_latents"])
        u_, c_ = eps.chunk(2)
        lat = sched.step(u_ + 7.5 * (c_ - u_), t, lat, return_dict=False)[0]
    with torch.no_grad():
        img = ts(lat)
    np.savez_compressed(os.path.join(GOLD, "sd15_loop_unipc%d_final.npz" % steps), latents=lat.numpy(), image=img.numpy(),
                        steps=np.int64(steps), guidance=np.float64(7.5))
    print(f"sd15 {steps}-step reference loop: {time.time() - t0:.0f}s, |lat|={lat.norm():.3f}, image range [{img.min():.2f},{img.max():.2f}]")


def golden_step_geometry(diffusers, name, images, hw, t=500, seed=0):
    """One reference step of the SD1.5-shaped nets at a geometry bench.py actually runs (config 2: 8 images = net batch 16
    at 64x64; config 5: 96x96 latents = 9216 tokens).  Fixture kept small: full noise_pred in float16-free fp32 only for the
    first and last sample of each CFG half plus norms and a strided sample of the whole tensor and of every tap."""
    from mirrorfusion_b200.config import SD15
    from mirrorfusion_b200.synth import make_inputs
    unet, bn, _, _ = build_reference_nets(diffusers, SD15, seed)
    inp = make_inputs(SD15, images, height=hw, width=hw)
    x = torch.cat([inp["latents"]] * 2)
    t0 = time.time()
    eps, d, m, u = ref_step(unet, bn, x, torch.tensor(t), inp["prompt_embeds"], inp["conditioning_latents"], 1.0)
    dt = time.time() - t0
    keep = sorted({0, images - 1, images, 2 * images - 1})
    out = {"t": np.int64(t), "images": np.int64(images), "hw": np.int64(hw), "seed": np.int64(seed),
           "noise_pred_l2": np.float64(eps.double().norm().item()),
           "noise_pred_l2_per_sample": eps.double().flatten(1).norm(dim=1).numpy(),
           "noise_pred_stride": np.int64(7), "noise_pred_strided": eps.flatten()[::7].numpy().copy(),
           "keep": np.array(keep, np.int64), "noise_pred_keep": eps[keep].numpy().copy()}
    for k, a in enumerate(list(d) + [m] + list(u)):
        out[f"tap{k:02d}_l2"] = np.float64(a.double().norm().item())
        out[f"tap{k:02d}_sample"] = a.flatten()[:: max(1, a.numel() // 4096)][:4096].numpy().copy()
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(f"{name}: {images} images at {hw}x{hw}: ref step {dt:.1f}s  |eps|={eps.norm():.4f}")


def golden_psnr50(diffusers, steps=50, name="sd15_loop_unipc50_sdvae.npz"):
    """north_star's final-image protocol at full size: the REFERENCE loop (SD1.5-shaped nets, 50 UniPC steps, CFG 7.5, fp32
    CPU) followed by the REFERENCE AutoencoderKL.decode with the SD VAE architecture at 512x512 (seeded weights from
    vae.make_vae_state_dict(SD_VAE), loaded strict), postprocessed to uint8 by the reference's VaeImageProcessor."""
    from diffusers.image_processor import VaeImageProcessor
    from mirrorfusion_b200.config import SD15
    from mirrorfusion_b200.synth import make_inputs
    from mirrorfusion_b200.vae import SD_VAE, make_vae_state_dict
    cfg = SD_VAE
    n = len(cfg.block_out_channels)
    vae = diffusers.AutoencoderKL(in_channels=3, out_channels=cfg.out_channels, down_block_types=("DownEncoderBlock2D",) * n,
                                  up_block_types=("UpDecoderBlock2D",) * n, block_out_channels=cfg.block_out_channels,
                                  layers_per_block=cfg.layers_per_block, latent_channels=cfg.latent_channels,
                                  norm_num_groups=cfg.norm_num_groups, sample_size=512, scaling_factor=cfg.scaling_factor).eval()
    vae.load_state_dict(make_vae_state_dict(cfg, 0, "both"), strict=True)
    unet, bn, _, _ = build_reference_nets(diffusers, SD15, 0)
    inp = make_inputs(SD15, 1)
    base = diffusers.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                   clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    sched = diffusers.UniPCMultistepScheduler.from_config(base.config)
    sched.set_timesteps(steps)
    lat = inp["latents"] * sched.init_noise_sigma
    t0 = time.time()
    traj = []
    for t in sched.timesteps:                                      # pipeline_brushnet.py:1250-1315
        x = torch.cat([lat] * 2)
        eps, *_ = ref_step(unet, bn, x, t, inp["prompt_embeds"], inp["conditioning_latents"])
        u_, c_ = eps.chunk(2)
        lat = sched.step(u_ + 7.5 * (c_ - u_), t, lat, return_dict=False)[0]
        traj.append(np.float64(lat.double().norm().item()))
    with torch.no_grad():
        img = vae.decode(lat / cfg.scaling_factor, return_dict=False)[0]                   # pipeline_brushnet.py:1342
    u8 = (VaeImageProcessor(vae_scale_factor=8).postprocess(img, output_type="np") * 255).round().astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, name), latents=lat.numpy(), image_u8=u8, latent_norms=np.array(traj),
                        image_f32_sample=img.flatten()[::64].numpy().copy(), steps=np.int64(steps), guidance=np.float64(7.5))
    print(f"{name}: {steps}-step reference loop + SD-VAE decode: {time.time() - t0:.0f}s, |lat|={lat.norm():.3f}, "
          f"image range [{img.min():.2f},{img.max():.2f}], uint8 mean {u8.mean():.1f}")


@torch.no_grad()
def golden_vae_decode(diffusers, name="tiny_vae_decode.npz", seed=0):
    """AutoencoderKL.decode of the reference on the seeded TINY_VAE decoder weights (vae.make_vae_state_dict)."""
    from mirrorfusion_b200.vae import TINY_VAE, make_vae_state_dict
    cfg = TINY_VAE
    n = len(cfg.block_out_channels)
    vae = diffusers.AutoencoderKL(in_channels=3, out_channels=cfg.out_channels, down_block_types=("DownEncoderBlock2D",) * n,
                                  up_block_types=("UpDecoderBlock2D",) * n, block_out_channels=cfg.block_out_channels,
                                  layers_per_block=cfg.layers_per_block, latent_channels=cfg.latent_channels,
                                  norm_num_groups=cfg.norm_num_groups, sample_size=32, scaling_factor=cfg.scaling_factor).eval()
    sd = make_vae_state_dict(cfg, seed, "both")
    vae.load_state_dict(sd, strict=True)                 # every encoder / decoder / quant tensor comes from the seeded dict
    z = torch.randn(2, cfg.latent_channels, 16, 16, generator=torch.Generator().manual_seed(4321)) * 3.0
    img = vae.decode(z).sample
    np.savez_compressed(os.path.join(GOLD, name), z=z.numpy(), image=img.numpy(), seed=seed)
    print(f"{name}: z {tuple(z.shape)} -> image {tuple(img.shape)}, |image| = {img.norm().item():.4f}")
    # encode: moments of the latent distribution for a synthetic image in [-1, 1], and one sample with known noise
    x = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(777)) * 2 - 1
    dist = vae.encode(x).latent_dist
    noise = torch.randn(dist.mean.shape, generator=torch.Generator().manual_seed(778))
    sample = dist.mean + dist.std * noise
    np.savez_compressed(os.path.join(GOLD, "tiny_vae_encode.npz"), x=x.numpy(), mean=dist.mean.numpy(), logvar=dist.logvar.numpy(),
                        noise=noise.numpy(), sample=sample.numpy(), seed=seed)
    print(f"tiny_vae_encode.npz: x {tuple(x.shape)} -> mean {tuple(dist.mean.shape)}, |mean| = {dist.mean.norm().item():.4f}")


def golden_checkpoint_layout(diffusers, name="micro_checkpoint_layout.json"):
    """What `save_pretrained(safe_serialization=True)` of the reference writes for the two networks (MICRO config):
    config.json verbatim + the safetensors header (names, dtypes, shapes).  The loader of mirrorfusion_b200.checkpoint is
    run on the reference-written directories right here, and must return exactly the seeded tensors."""
    import json
    import tempfile
    from safetensors import safe_open
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200 import checkpoint as CK
    unet, bn, usd, bsd = build_reference_nets(diffusers, MICRO)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for net, mod, sd in (("unet", unet, usd), ("brushnet", bn, bsd)):
            d = os.path.join(tmp, net)
            mod.save_pretrained(d, safe_serialization=True)
            with safe_open(os.path.join(d, CK.WEIGHTS[0]), framework="pt") as sf:
                tensors = {k: [str(sf.get_slice(k).get_dtype()), list(sf.get_slice(k).get_shape())] for k in sf.keys()}
            out[net] = {"files": sorted(os.listdir(d)), "config": json.load(open(os.path.join(d, "config.json"))), "tensors": tensors}
        cfg, lu, lb = CK.load_mirrorfusion(os.path.join(tmp, "unet"), os.path.join(tmp, "brushnet"))
        import dataclasses
        assert dataclasses.asdict(cfg) == dataclasses.asdict(MICRO), (cfg, MICRO)
        assert all(torch.equal(lu[k], usd[k]) for k in usd) and all(torch.equal(lb[k], bsd[k]) for k in bsd)
        out["loader_verified_on_reference_written_files"] = True
    with open(os.path.join(GOLD, name), "w") as f:
        json.dump(out, f, indent=0)
    print(f"{name}: {len(out['unet']['tensors'])} + {len(out['brushnet']['tensors'])} tensors; loader verified on the reference's files")


def golden_prep(diffusers, name="prep_golden.npz"):
    """Input / output processing by the reference's OWN functions: VaeImageProcessor.preprocess on PIL images (as
    E/test_brushnet.py passes them), the mask rule + F.interpolate of pipeline_brushnet.py:1139,1190-1202,
    HDF5Dataset.apply_transforms_depth (its source is exec'd standalone: the module itself imports h5py, which is not
    installed), VaeImageProcessor.postprocess."""
    import ast
    import torch.nn.functional as F
    from PIL import Image
    from torchvision import transforms
    from diffusers.image_processor import VaeImageProcessor
    rng = np.random.default_rng(5)
    N, S, f = 2, 64, 8
    rgb = rng.integers(0, 256, (N, S, S, 3), dtype=np.uint8)
    mask = np.zeros((N, S, S), np.uint8)
    mask[0, 10:40, 20:50] = 255
    mask[1, 5:30, 3:33] = 255
    depth = (rng.random((N, S, S), dtype=np.float32) * 6.0 + 0.2).astype(np.float32)
    proc = VaeImageProcessor(vae_scale_factor=f, do_convert_rgb=True)                     # pipeline_brushnet.py:230
    img_t = torch.cat([proc.preprocess(Image.fromarray(rgb[i]), height=S, width=S) for i in range(N)])
    m_t = torch.cat([proc.preprocess(Image.fromarray(mask[i]).convert("RGB"), height=S, width=S) for i in range(N)])
    m1 = (m_t.sum(1)[:, None] < 0).to(img_t.dtype)                                         # :1139
    m_lat = F.interpolate(m1, size=(S // f, S // f))                                       # :1190-1196
    src = open("/root/reference/MirrorFusion/examples/brushnet/dataset/dataset.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "apply_transforms_depth")
    fn.decorator_list = []
    ns = {"np": np, "torch": torch, "transforms": transforms}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "dataset.py", "exec"), ns)
    d_t = torch.stack([ns["apply_transforms_depth"](depth[i], mask[i], resolution=S) for i in range(N)])   # [N,1,S,S]
    d_lat = F.interpolate(d_t, size=(S // f, S // f))                                      # :1198-1202
    dec = torch.randn(N, 3, S, S, generator=torch.Generator().manual_seed(3)) * 0.8
    out_u8 = (proc.postprocess(dec, output_type="np") * 255).round().astype(np.uint8)      # numpy_to_pil's rounding
    np.savez_compressed(os.path.join(GOLD, name), rgb=rgb, mask=mask, depth=depth, image=img_t.numpy(), mask_lat=m_lat.numpy(),
                        depth_lat=d_lat.numpy(), decoded=dec.numpy(), out_u8=out_u8, factor=f)
    print(f"{name}: image {tuple(img_t.shape)}, mask_lat {tuple(m_lat.shape)} (sum {m_lat.sum().item():.0f}), depth_lat {tuple(d_lat.shape)}")


def golden_signatures(diffusers, name):
    """Parameter names (in order) of the reference entry points the drop-in classes mirror."""
    import inspect
    import json
    from diffusers.models.attention_processor import AttnProcessor2_0
    targets = {
        "BrushNetModel.forward": diffusers.BrushNetModel.forward,
        "UNet2DConditionModel.forward": diffusers.UNet2DConditionModel.forward,
        "StableDiffusionBrushNetPipeline.__call__": diffusers.StableDiffusionBrushNetPipeline.__call__,
        "AttnProcessor2_0.__call__": AttnProcessor2_0.__call__,
        "UniPCMultistepScheduler.step": diffusers.UniPCMultistepScheduler.step,
        "DDIMScheduler.step": diffusers.DDIMScheduler.step,
        "UniPCMultistepScheduler.set_timesteps": diffusers.UniPCMultistepScheduler.set_timesteps,
        "UniPCMultistepScheduler.scale_model_input": diffusers.UniPCMultistepScheduler.scale_model_input,
    }
    out = {}
    for k, fn in targets.items():
        sig = inspect.signature(fn)
        out[k] = [{"name": p.name, "default": None if p.default is inspect._empty else repr(p.default),
                   "kind": str(p.kind)} for p in sig.parameters.values()]
    with open(os.path.join(GOLD, name), "w") as f:
        json.dump(out, f, indent=1)
    print(f"{name}: {len(out)} reference signatures")


def main():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("reference not mounted at /root/reference; golden vectors can only be made in the build container")
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    diffusers = import_reference()
    from mirrorfusion_b200.config import MICRO, TINY, SD15
    which = sys.argv[1:] or ["sched", "micro", "tiny", "sd15", "sigs"]
    if "psnr" in which:
        golden_psnr(diffusers)
    if "psnr50" in which:
        golden_psnr50(diffusers)
    if "b16" in which:
        golden_step_geometry(diffusers, "sd15_step_b16.npz", images=8, hw=64)
    if "g96" in which:
        golden_step_geometry(diffusers, "sd15_step_96.npz", images=2, hw=96)
    if "vae" in which:
        golden_vae_decode(diffusers)
    if "prep" in which:
        golden_prep(diffusers)
    if "ckpt" in which:
        golden_checkpoint_layout(diffusers)
    if "sigs" in which:
        golden_signatures(diffusers, "reference_signatures.json")
    if "sched" in which:
        golden_sched_only(diffusers, "sched_traj.npz")
    if "guess" in which:
        golden_step_guess_mode(diffusers, TINY, "tiny_step_guess_mode.npz")
    if "sched_eta" in which:
        golden_sched_eta(diffusers, "sched_eta_traj.npz")
    if "micro" in which:
        golden_step(diffusers, MICRO, "micro_step.npz", images=2, t=321, scale=0.8)
        golden_loop(diffusers, MICRO, "micro_loop_ddim4.npz", "ddim", 4)
        golden_loop(diffusers, MICRO, "micro_loop_unipc6.npz", "unipc", 6)
    if "tiny" in which:
        golden_step(diffusers, TINY, "tiny_step.npz", images=1, t=500)
        golden_loop(diffusers, TINY, "tiny_loop_unipc8.npz", "unipc", 8)
    if "sd15" in which:
        golden_step(diffusers, SD15, "sd15_step.npz", images=1, t=500, full_taps=False)


if __name__ == "__main__":
    main()
