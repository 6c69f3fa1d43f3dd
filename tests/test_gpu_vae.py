"""AutoencoderKL.decode on the libmfb200 kernels (mirrorfusion_b200/vae.py) against the reference's own output
(tests/golden/tiny_vae_decode.npz, made by oracle/make_golden.py from the reference AutoencoderKL) and against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mirrorfusion_b200 import ops
from mirrorfusion_b200.vae import SD_VAE, TINY_VAE, VaeConfig, VaeDecoderEngine, make_vae_state_dict
from test_gpu_model import record, rel


def psnr_u8(a, b):
    """PSNR of the uint8 images ((x / 2 + 0.5).clamp(0, 1) * 255, VaeImageProcessor.postprocess) — the protocol of
    M/metrics/metrics.py:62-67,197-200."""
    q = lambda t: ((torch.as_tensor(t).float().cpu() / 2 + 0.5).clamp(0, 1) * 255).round()
    mse = ((q(a) - q(b)) ** 2).mean().item()
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("fp32", 1e-4)])
def test_tiny_vae_decode_vs_reference_golden(golden_dir, precision, tol):
    g = np.load(os.path.join(golden_dir, "tiny_vae_decode.npz"))
    sd = make_vae_state_dict(TINY_VAE, int(g["seed"]))
    z = torch.from_numpy(g["z"])
    with ops.precision(precision):
        eng = VaeDecoderEngine(TINY_VAE, sd, z.shape[0], z.shape[2], z.shape[3])
    img = eng.decode(z.cuda())
    e = rel(img, g["image"])
    p = psnr_u8(img, g["image"])
    record("tiny_vae_decode_vs_reference", precision=precision, image_rel_l2=e, psnr_db=p)
    assert e < tol
    assert p >= 40.0          # north_star: final decoded images within PSNR >= 40 dB of the reference


def test_vae_decode_gemm_attention_path_vs_oracle():
    """A 512-channel mid block (the SD VAE's): the single head of dim 512 runs as two GEMMs around the row softmax.
    Two-level decoder (256, 512), 16x16 latents, against the oracle on the host."""
    from oracle.vae_oracle import vae_decode
    cfg = VaeConfig(block_out_channels=(256, 512), layers_per_block=1)
    sd = make_vae_state_dict(cfg, 3)
    z = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(9)) * 3.0
    eng = VaeDecoderEngine(cfg, sd, 2, 16, 16)
    assert any(n is not None and "N=256 K=512" in n for n in eng.notes)      # S = Q K^T as a GEMM: the d = 512 path is taken
    img = eng.decode(z.cuda())
    ref = vae_decode(sd, cfg, z)
    e, p = rel(img, ref), psnr_u8(img, ref)
    record("vae_decode_d512_gemm_attention_vs_oracle", image_rel_l2=e, psnr_db=p)
    assert e < 2e-2 and p >= 40.0


def test_sd_vae_decode_512px_runs_and_matches_fp32_mode_on_one_image():
    """The real geometry (SD VAE, 64x64 latents -> 512x512): bf16 product path against the fp32 parity mode of the
    same program (the reference cannot travel to the GPU box; the fp32 mode is pinned to it at 1e-4 above)."""
    cfg = SD_VAE
    sd = make_vae_state_dict(cfg, 0)
    z = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(5)) * 4.0
    eng = VaeDecoderEngine(cfg, sd, 1, 64, 64)
    img = eng.decode(z.cuda()).clone()
    assert tuple(img.shape) == (1, 3, 512, 512)
    with ops.precision("fp32"):
        e32 = VaeDecoderEngine(cfg, sd, 1, 64, 64)
    ref = e32.decode(z.cuda())
    e, p = rel(img, ref), psnr_u8(img, ref)
    record("sd_vae_decode_512px_bf16_vs_fp32_mode", image_rel_l2=e, psnr_db=p)
    assert e < 2e-2 and p >= 40.0


def test_pipeline_call_decodes_images_on_the_kernels():
    """MirrorFusionB200Pipeline(vae_state_dict=...) with output_type='pt': denoise loop + VAE decode, both on the
    kernels, against the oracle's loop followed by the oracle's decode (pipeline_brushnet.py:1337-1342)."""
    from mirrorfusion_b200 import pipeline as P
    from mirrorfusion_b200.config import TINY
    from mirrorfusion_b200.synth import make_inputs, make_state_dict
    from oracle import mf_oracle as O
    from oracle.vae_oracle import vae_decode
    cfg = TINY
    usd, bsd = make_state_dict(cfg, "unet"), make_state_dict(cfg, "brushnet")
    vsd = make_vae_state_dict(TINY_VAE, 0)
    inp = make_inputs(cfg, 1)
    pipe = P.MirrorFusionB200Pipeline(usd, bsd, scheduler=P.B200DDIMScheduler(), cfg=cfg, vae_state_dict=vsd, vae_cfg=TINY_VAE)
    img = pipe(prompt_embeds=inp["prompt_embeds"][1:].cuda(), negative_prompt_embeds=inp["prompt_embeds"][:1].cuda(),
               conditioning_latents=inp["conditioning_latents"][:1].cuda(), latents=inp["latents"].cuda(),
               num_inference_steps=4, guidance_scale=7.5, output_type="pt").images
    s = cfg.sample_size * 2 ** (len(TINY_VAE.block_out_channels) - 1)
    assert tuple(img.shape) == (1, 3, s, s)
    with torch.no_grad():
        lat = O.denoise_loop(usd, bsd, cfg, O.DDIMOracle(), inp["latents"], inp["prompt_embeds"], inp["conditioning_latents"], 4, 7.5)
    ref = vae_decode(vsd, TINY_VAE, lat / 0.18215)
    p = psnr_u8(img, ref)
    record("tiny_pipeline_images_vs_oracle", psnr_db=p, image_rel_l2=rel(img, ref))
    assert p >= 40.0


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("fp32", 1e-4)])
def test_tiny_vae_encode_vs_reference_golden(golden_dir, precision, tol):
    """AutoencoderKL.encode on the kernels (VaeEncoderEngine: pad0 stride-2 convs, conv_out . quant_conv folded, fp32 moments)
    against the reference's own latent distribution for the same seeded weights and image."""
    from mirrorfusion_b200.vae import VaeEncoderEngine
    g = np.load(os.path.join(golden_dir, "tiny_vae_encode.npz"))
    sd = make_vae_state_dict(TINY_VAE, int(g["seed"]), "both")
    x = torch.from_numpy(g["x"])
    with ops.precision(precision):
        eng = VaeEncoderEngine(TINY_VAE, sd, x.shape[0], x.shape[2], x.shape[3])
    lat = eng.encode(x.cuda(), noise=torch.from_numpy(g["noise"]).cuda(), scale=1.0)
    em, ev, es = rel(eng.mean, g["mean"]), rel(eng.logvar, g["logvar"]), rel(lat, g["sample"])
    record("tiny_vae_encode_vs_reference", precision=precision, mean=em, logvar=ev, sample=es)
    assert em < tol and ev < tol and es < tol
    mode = eng.encode(x.cuda(), noise=None, scale=TINY_VAE.scaling_factor).cpu()
    assert rel(mode, g["mean"] * TINY_VAE.scaling_factor) < tol


def test_sd_vae_encode_512px_bf16_vs_fp32_mode():
    """The real geometry (SD VAE, 512x512 -> 64x64 latents, mid-block attention over 4096 tokens with d = 512)."""
    from mirrorfusion_b200.vae import VaeEncoderEngine
    sd = make_vae_state_dict(SD_VAE, 0, "encoder")
    x = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(6)) * 2 - 1
    eng = VaeEncoderEngine(SD_VAE, sd, 1, 512, 512)
    eng.encode(x.cuda())
    mean = eng.mean.clone()
    assert tuple(mean.shape) == (1, 4, 64, 64)
    with ops.precision("fp32"):
        e32 = VaeEncoderEngine(SD_VAE, sd, 1, 512, 512)
    e32.encode(x.cuda())
    e = rel(mean, e32.mean)
    record("sd_vae_encode_512px_bf16_vs_fp32_mode", mean=e)
    assert e < 2e-2
