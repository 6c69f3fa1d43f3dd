"""Pins oracle/mf_oracle.py against vectors produced by the reference itself
(oracle/make_golden.py) and against the reference's own known-answer tests."""
import os

import numpy as np
import pytest
import torch

from mirrorfusion_b200.config import MICRO, TINY, SD15, param_shapes, tap_channels
from mirrorfusion_b200.synth import make_state_dict, make_inputs
from oracle import mf_oracle as O


def _load(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated")
    return np.load(p)


def rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def test_param_census_matches_survey():
    # 859 520 964 UNet params, 618 832 960 BrushNet params; 686 / 322 tensors
    u = param_shapes(SD15, "unet")
    b = param_shapes(SD15, "brushnet")
    assert len(u) == 686 and len(b) == 322
    assert sum(int(np.prod(s)) for _, s in u) == 859_520_964
    assert sum(int(np.prod(s)) for _, s in b) == 618_832_960
    d, m, up = tap_channels(SD15)
    assert len(d) == 12 and len(up) == 15 and m == 1280


def test_timestep_embedding_reference_literals():
    # T/models/test_layers_utils.py:94-120 exercises get_timestep_embedding; with flip_sin_to_cos the
    # layout is [cos || sin]; spot values follow from the closed form.
    t = torch.tensor([0.0, 7.0])
    e = O.timestep_embedding(t, 320)
    assert e.shape == (2, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))
    assert abs(e[1, 0].item() - np.cos(7.0)) < 1e-6 and abs(e[1, 160].item() - np.sin(7.0)) < 1e-6
    f1 = np.exp(-np.log(10000.0) * 1 / 160)
    assert abs(e[1, 161].item() - np.sin(7.0 * f1)) < 1e-6


@pytest.mark.parametrize("cfg,name", [(MICRO, "micro_step.npz"), (TINY, "tiny_step.npz")])
def test_step_matches_reference(golden_dir, cfg, name):
    g = _load(golden_dir, name)
    images = int(g["images"])
    usd = make_state_dict(cfg, "unet", int(g["seed"]))
    bsd = make_state_dict(cfg, "brushnet", int(g["seed"]))
    inp = make_inputs(cfg, images)
    x = torch.cat([inp["latents"]] * 2)
    with torch.no_grad():
        eps, (d, m, u) = O.noise_pred_step(usd, bsd, cfg, x, torch.tensor(int(g["t"])), inp["prompt_embeds"],
                                           inp["conditioning_latents"], float(g["scale"]))
        plain = O.unet_forward(usd, cfg, x, torch.tensor(int(g["t"])), inp["prompt_embeds"])
    taps = list(d) + [m] + list(u)
    for k, a in enumerate(taps):
        assert rel(a, g[f"tap{k:02d}"]) < 2e-5, f"tap {k}"
    assert rel(eps, g["noise_pred"]) < 1e-4   # fp32 summation-order noise (SDPA vs matmul+softmax)
    assert rel(plain, g["noise_pred_no_taps"]) < 1e-4


@pytest.mark.parametrize("name,kind,steps", [("micro_loop_ddim4.npz", "ddim", 4), ("micro_loop_unipc6.npz", "unipc", 6)])
def test_loop_matches_reference(golden_dir, name, kind, steps):
    g = _load(golden_dir, name)
    cfg = MICRO
    usd = make_state_dict(cfg, "unet", 0)
    bsd = make_state_dict(cfg, "brushnet", 0)
    inp = make_inputs(cfg, 1)
    sched = O.DDIMOracle() if kind == "ddim" else O.UniPCOracle()
    with torch.no_grad():
        lat, trace = O.denoise_loop(usd, bsd, cfg, sched, inp["latents"], inp["prompt_embeds"],
                                    inp["conditioning_latents"], steps, 7.5, 1.0, return_trace=True)
    assert np.array_equal(sched.timesteps.numpy(), g["timesteps"])
    for i, (eps, l) in enumerate(trace):
        assert rel(eps, g["noise_pred"][i]) < 1e-4, f"eps step {i}"
        assert rel(l, g["latents"][i]) < 1e-4, f"latents step {i}"


@pytest.mark.parametrize("kind,n", [("ddim", 4), ("ddim", 10), ("unipc", 5), ("unipc", 10), ("unipc", 50)])
def test_scheduler_trajectories(golden_dir, kind, n):
    g = _load(golden_dir, "sched_traj.npz")
    s = O.DDIMOracle() if kind == "ddim" else O.UniPCOracle()
    s.set_timesteps(n)
    assert np.array_equal(s.timesteps.numpy(), g[f"{kind}{n}_timesteps"])
    if kind == "unipc":
        assert np.allclose(s.sigmas.numpy(), g[f"{kind}{n}_sigmas"], rtol=0, atol=0)
    x = torch.from_numpy(g["x0"]).clone()
    for i, t in enumerate(s.timesteps):
        eps = torch.sin(3.0 * x + 0.01 * float(t)) * 0.9 + 0.1 * x
        x = s.step(eps, t, x)
        assert rel(x, g[f"{kind}{n}_traj"][i]) < 2e-6, f"step {i}"


def test_ddim_known_answer_from_reference_tests():
    # T/schedulers/test_scheduler_ddim.py:114-121 (test_full_loop_no_noise): linear betas 1e-4..0.02,
    # clip_sample=True, 10 steps, dummy model/sample of T/schedulers/test_schedulers.py:326-368
    # -> sum |x| = 172.0067, mean |x| = 0.223967
    s = O.DDIMOracle(beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True,
                     set_alpha_to_one=True, steps_offset=0)
    s.set_timesteps(10)
    num_elems = 4 * 3 * 8 * 8
    sample = (torch.arange(num_elems).reshape(4, 3, 8, 8) / num_elems).permute(3, 0, 1, 2)
    for t in s.timesteps:
        residual = sample * int(t) / (int(t) + 1)
        sample = s.step(residual, t, sample)
    assert abs(sample.abs().sum().item() - 172.0067) < 1e-2
    assert abs(sample.abs().mean().item() - 0.223967) < 1e-3


def test_unipc_known_answer_from_reference_tests():
    # T/schedulers/test_scheduler_unipc.py:206-210 (test_full_loop_no_noise): linear betas, solver_order 2,
    # bh2, 10 steps, dummy model -> mean |x| = 0.2464
    s = O.UniPCOracle(beta_start=0.0001, beta_end=0.02, beta_schedule="linear")
    s.set_timesteps(10)
    num_elems = 4 * 3 * 8 * 8
    sample = (torch.arange(num_elems).reshape(4, 3, 8, 8) / num_elems).permute(3, 0, 1, 2)
    for t in s.timesteps:
        residual = sample * int(t) / (int(t) + 1)
        sample = s.step(residual, t, sample)
    assert abs(sample.abs().mean().item() - 0.2464) < 1e-3


def test_vae_decode_oracle_matches_reference_golden(golden_dir):
    """oracle/vae_oracle.py against AutoencoderKL.decode of the reference itself on the seeded TINY_VAE weights
    (the VAE decode is the next row after the denoise step)."""
    from mirrorfusion_b200.vae import TINY_VAE, make_vae_state_dict, vae_decoder_param_shapes, SD_VAE
    from oracle.vae_oracle import vae_decode
    g = _load(golden_dir, "tiny_vae_decode.npz")
    sd = make_vae_state_dict(TINY_VAE, int(g["seed"]))
    img = vae_decode(sd, TINY_VAE, torch.from_numpy(g["z"]))
    assert tuple(img.shape) == g["image"].shape
    assert rel(img, g["image"]) < 1e-5
    # decoder + post_quant_conv of the SD VAE: 49 490 199 parameters (AutoencoderKL(block_out_channels=(128,256,512,512)))
    assert sum(int(np.prod(s)) for _, s in vae_decoder_param_shapes(SD_VAE)) == 49_490_199


def test_vae_encode_oracle_matches_reference_golden(golden_dir):
    from mirrorfusion_b200.vae import TINY_VAE, make_vae_state_dict, vae_encoder_param_shapes, SD_VAE
    from oracle.vae_oracle import vae_encode_moments, latent_sample
    g = _load(golden_dir, "tiny_vae_encode.npz")
    sd = make_vae_state_dict(TINY_VAE, int(g["seed"]), "both")
    mean, logvar = vae_encode_moments(sd, TINY_VAE, torch.from_numpy(g["x"]))
    assert rel(mean, g["mean"]) < 1e-5 and rel(logvar, g["logvar"]) < 1e-5
    assert rel(latent_sample(mean, logvar, torch.from_numpy(g["noise"])), g["sample"]) < 1e-5
    # encoder + quant_conv of the SD VAE: 34 163 592 + 72 parameters
    assert sum(int(np.prod(s)) for _, s in vae_encoder_param_shapes(SD_VAE)) == 34_163_664


def test_prep_oracle_matches_reference_functions(golden_dir):
    """oracle/prep_oracle.py against VaeImageProcessor.preprocess / postprocess, the mask rule + F.interpolate of the pipeline
    and HDF5Dataset.apply_transforms_depth, all run by oracle/make_golden.py from the reference's own code."""
    from oracle import prep_oracle as PO
    g = _load(golden_dir, "prep_golden.npz")
    f = int(g["factor"])
    assert np.abs(PO.prep_image(g["rgb"]) - g["image"]).max() < 1e-6
    assert np.array_equal(PO.prep_mask(g["mask"], f), g["mask_lat"])
    assert np.abs(PO.prep_depth(g["depth"], g["mask"], f) - g["depth_lat"]).max() < 1e-6
    assert np.array_equal(PO.post_image(g["decoded"]), g["out_u8"])


def test_resize_oracle_vs_torchvision():
    """oracle/resize_oracle.py against the transform the reference applies (E/dataset/dataset.py:70-76,86-92,155-165):
    torchvision `Resize(res, BICUBIC)` + `CenterCrop(res)` on float tensors."""
    tv = pytest.importorskip("torchvision")
    from torchvision import transforms
    from oracle.resize_oracle import crop_offsets, resize_crop_bicubic, resized_size
    rng = np.random.default_rng(0)
    for Hs, Ws, res in [(512, 512, 512), (384, 512, 256), (512, 384, 256), (300, 451, 256), (150, 200, 256), (97, 131, 40), (33, 33, 64)]:
        x = rng.standard_normal((2, Hs, Ws)).astype(np.float32)
        t = transforms.Compose([transforms.Resize(res, interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(res)])
        want = t(torch.from_numpy(x)).numpy()
        got = resize_crop_bicubic(x, res)
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-5, (Hs, Ws, res)
        if res % 8 == 0:
            assert np.abs(resize_crop_bicubic(x, res, 8) - want[:, ::8, ::8]).max() < 1e-5
    assert resized_size(300, 451, 256) == (256, 384) and crop_offsets(256, 385, 256) == (0, 64)      # round half to even: 64.5 -> 64


def test_guess_mode_step_matches_reference(golden_dir):
    """guess_mode with CFG (pipeline_brushnet.py:1262-1301; brushnet.py:896-902): BrushNet on the conditional half with log-spaced tap
    scales, zeros for the unconditional half — vector written by the reference's own BrushNetModel.forward(guess_mode=True) + UNet."""
    g = _load(golden_dir, "tiny_step_guess_mode.npz")
    cfg, images = TINY, int(g["images"])
    usd, bsd = make_state_dict(cfg, "unet", int(g["seed"])), make_state_dict(cfg, "brushnet", int(g["seed"]))
    inp = make_inputs(cfg, images)
    x = torch.cat([inp["latents"]] * 2)
    with torch.no_grad():
        eps, (d, m, u) = O.noise_pred_step(usd, bsd, cfg, x, torch.tensor(int(g["t"])), inp["prompt_embeds"],
                                           inp["conditioning_latents"][images:], float(g["scale"]), guess_mode=True)
    assert rel(eps, g["noise_pred"]) < 1e-4
    for k, a in enumerate(list(d) + [m] + list(u)):
        assert abs(a.double().norm().item() / float(g[f"tap{k:02d}_l2"]) - 1) < 1e-4
        assert float(a[:images].abs().max()) == 0.0                     # the unconditional half gets zeros
