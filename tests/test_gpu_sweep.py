"""Batched eval sweep (mirrorfusion_b200/sweep.py): kernels for the pre/post-processing against vectors produced by the
reference's own functions, the whole sweep against the oracle chain per item, and invariance to batching / sharding."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mirrorfusion_b200 import ops
from mirrorfusion_b200.config import TINY
from mirrorfusion_b200.synth import make_state_dict
from mirrorfusion_b200.vae import TINY_VAE, make_vae_state_dict
from test_gpu_model import record


def test_prep_and_post_kernels_vs_reference_functions(golden_dir):
    g = np.load(os.path.join(golden_dir, "prep_golden.npz"))
    f = int(g["factor"])
    N, S = g["rgb"].shape[0], g["rgb"].shape[1]
    ops.lib()
    dev = "cuda"
    img = torch.empty(N, 3, S, S, device=dev)
    ops.prep_image_u8(torch.from_numpy(g["rgb"]).to(dev), img)
    assert (img.cpu() - torch.from_numpy(g["image"])).abs().max().item() < 1e-6
    ml, dl = torch.empty(N, 1, S // f, S // f, device=dev), torch.empty(N, 1, S // f, S // f, device=dev)
    ops.prep_mask_depth(torch.from_numpy(g["mask"]).to(dev), torch.from_numpy(g["depth"]).to(dev), ml, dl,
                        torch.zeros(N, dtype=torch.int32, device=dev), factor=f)
    assert torch.equal(ml.cpu(), torch.from_numpy(g["mask_lat"]))
    assert (dl.cpu() - torch.from_numpy(g["depth_lat"])).abs().max().item() < 1e-6
    u8 = torch.empty(N, S, S, 3, dtype=torch.uint8, device=dev)
    ops.post_image_u8(torch.from_numpy(g["decoded"]).to(dev), u8)
    assert np.array_equal(u8.cpu().numpy(), g["out_u8"])


RESIZE_CASES = [(512, 512, 512), (768, 1024, 512), (1024, 768, 512), (600, 901, 512), (300, 400, 512), (97, 131, 40), (33, 33, 64)]


@pytest.mark.parametrize("Hs,Ws,res", RESIZE_CASES)
def test_resize_crop_kernel_vs_oracle_and_torchvision(Hs, Ws, res):
    """mfb_resize_crop_bicubic = transforms.Resize(res, BICUBIC) + CenterCrop(res) on float tensors (E/dataset/dataset.py:70-76,
    86-92,155-165): down- and up-scaling, portrait / landscape / odd sizes; step = 8 is the crop sampled at latent resolution."""
    from oracle.resize_oracle import resize_crop_bicubic
    rng = np.random.default_rng(Hs * 7 + Ws)
    x = rng.standard_normal((3, Hs, Ws)).astype(np.float32)
    want = resize_crop_bicubic(x, res)
    ops.lib()
    xd = torch.from_numpy(x).cuda()
    out = torch.empty(3, res, res, device="cuda")
    ops.resize_crop_bicubic(xd, out, res)
    assert np.abs(out.cpu().numpy() - want).max() < 1e-5
    try:
        from torchvision import transforms
    except ImportError:
        transforms = None
    if transforms is not None:                 # the reference's own transform, where torchvision is installed
        t = transforms.Compose([transforms.Resize(res, interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(res)])
        assert (out.cpu() - t(torch.from_numpy(x))).abs().max().item() < 1e-5
    if res % 8 == 0:
        o8 = torch.empty(3, res // 8, res // 8, device="cuda")
        ops.resize_crop_bicubic(xd, o8, res, step=8)
        assert np.abs(o8.cpu().numpy() - want[:, ::8, ::8]).max() < 1e-5


def test_depth_normalize_kernel_vs_oracle():
    from oracle import prep_oracle as PO
    rng = np.random.default_rng(5)
    N, H, W = 3, 40, 56
    depth = (rng.random((N, H, W), dtype=np.float32) * 6 - 0.5).astype(np.float32)           # some negative: clipped to 0
    mask = np.zeros((N, H, W), np.uint8)
    for i in range(N):
        mask[i, 3 + i:20, 5:30 + i] = 255
    ops.lib()
    out = torch.empty(N, H, W, device="cuda")
    ops.depth_normalize(torch.from_numpy(depth).cuda(), torch.from_numpy(mask).cuda(), out, torch.zeros(N, dtype=torch.int32, device="cuda"))
    assert np.abs(out.cpu().numpy() - PO.prep_depth(depth, mask, 1)[:, 0]).max() < 1e-6


def _inputs(S, px, cfg, seed=0):
    rng = np.random.default_rng(seed)
    rgb = rng.integers(0, 256, (S, px, px, 3), dtype=np.uint8)
    mask = np.zeros((S, px, px), np.uint8)
    for i in range(S):
        mask[i, 4 + i:18 + i, 6:22 + 2 * i] = 255
    rgb[mask > 0] = 0                                        # masked image: black inside the mirror region
    depth = (rng.random((S, px, px), dtype=np.float32) * 4 + 0.5).astype(np.float32)
    g = torch.Generator().manual_seed(seed + 1)
    pe = torch.randn(S, 77, cfg.cross_attention_dim, generator=g)
    ne = torch.randn(S, 77, cfg.cross_attention_dim, generator=g)
    return rgb, mask, depth, pe, ne


def _sweep(images_per_call, repeats=2, steps=3):
    from mirrorfusion_b200.pipeline import B200DDIMScheduler
    from mirrorfusion_b200.sweep import EvalSweep
    cfg = TINY
    px = cfg.sample_size * 2                                 # TINY_VAE has one downsampling level
    return EvalSweep(cfg, make_state_dict(cfg, "unet"), make_state_dict(cfg, "brushnet"), TINY_VAE,
                     make_vae_state_dict(TINY_VAE, 0, "both"), B200DDIMScheduler, H=px, W=px, images_per_call=images_per_call,
                     repeats=repeats, num_inference_steps=steps), px


def test_sweep_vs_oracle_chain_per_item():
    """Every stage on the kernels vs the oracle chain (prep -> VAE encode + sample -> conditioning -> loop -> VAE decode ->
    uint8) evaluated item by item like the reference visits them."""
    from mirrorfusion_b200.sweep import item_generator
    from oracle import mf_oracle as O
    from oracle import prep_oracle as PO
    from oracle.vae_oracle import latent_sample, vae_decode, vae_encode_moments
    sw, px = _sweep(images_per_call=4)
    cfg, S, seed = TINY, 3, 11
    rgb, mask, depth, pe, ne = _inputs(S, px, cfg)
    got, items = sw.run(rgb, mask, depth, pe, ne, seed=seed)
    assert items == [(i, k) for i in range(S) for k in range(2)] and got.shape == (6, px, px, 3)
    usd, bsd, vsd = make_state_dict(cfg, "unet"), make_state_dict(cfg, "brushnet"), make_vae_state_dict(TINY_VAE, 0, "both")
    worst = 1e9
    for j, (i, k) in enumerate(items):
        img = torch.from_numpy(PO.prep_image(rgb[i:i + 1]))
        mean, logvar = vae_encode_moments(vsd, TINY_VAE, img)
        vn = torch.randn(4, sw.h, sw.w, generator=item_generator(seed, j, 1))[None]
        lat = latent_sample(mean, logvar, vn) * TINY_VAE.scaling_factor
        cond = torch.cat([lat, torch.from_numpy(PO.prep_mask(mask[i:i + 1], sw.f)), torch.from_numpy(PO.prep_depth(depth[i:i + 1], mask[i:i + 1], sw.f))], 1)
        x0 = torch.randn(4, sw.h, sw.w, generator=item_generator(seed, j, 0))[None]
        with torch.no_grad():
            x = O.denoise_loop(usd, bsd, cfg, O.DDIMOracle(), x0, torch.cat([ne[i:i + 1], pe[i:i + 1]]), torch.cat([cond, cond]), 3, 7.5)
        ref = PO.post_image(vae_decode(vsd, TINY_VAE, x / TINY_VAE.scaling_factor).numpy())[0]
        mse = ((got[j].astype(np.float64) - ref.astype(np.float64)) ** 2).mean()
        worst = min(worst, 10 * np.log10(255.0 ** 2 / max(mse, 1e-12)))
    record("sweep_vs_oracle_chain", worst_psnr_db=worst, items=len(items))
    assert worst >= 40.0          # north_star: decoded images within PSNR >= 40 dB


def test_sweep_is_invariant_to_batching_and_sharding():
    """An item's image depends on (seed, global item index) only: the same bytes whether the sweep runs as one rank with
    4 items per call or as two ranks with 2 items per call (BASELINE config 3: any GPU count reproduces the outputs)."""
    cfg, S, seed = TINY, 3, 5
    sw4, px = _sweep(images_per_call=4)
    rgb, mask, depth, pe, ne = _inputs(S, px, cfg, seed=2)
    whole, items = sw4.run(rgb, mask, depth, pe, ne, seed=seed)
    del sw4
    sw2, _ = _sweep(images_per_call=2)
    parts = [sw2.run(rgb, mask, depth, pe, ne, seed=seed, rank=r, world=2) for r in range(2)]
    assert parts[0][1] + parts[1][1] == items
    assert np.array_equal(np.concatenate([parts[0][0], parts[1][0]]), whole)


def test_sweep_depth_at_its_own_resolution():
    """Depth maps that are not at the target resolution go through apply_transforms_depth (normalise at their own size, bicubic
    antialiased resize + centre crop, E/dataset/dataset.py:98-166) fused with the nearest sampling to latent size."""
    from oracle import prep_oracle as PO
    from oracle.resize_oracle import resize_crop_bicubic
    sw, px = _sweep(images_per_call=4)
    cfg, S = TINY, 2
    rgb, mask, _, pe, ne = _inputs(S, px, cfg)
    rng = np.random.default_rng(3)
    Hd, Wd = 3 * px // 2 + 1, 2 * px
    depth = (rng.random((S, Hd, Wd), dtype=np.float32) * 4 + 0.5).astype(np.float32)
    dmask = np.zeros((S, Hd, Wd), np.uint8)
    dmask[:, 5:30, 8:40] = 255
    with pytest.raises(ValueError):
        sw.run(rgb, mask, depth, pe, ne, seed=1)
    got, items = sw.run(rgb, mask, depth, pe, ne, seed=1, depth_mask=dmask)
    assert got.shape == (4, px, px, 3) and items == [(0, 0), (0, 1), (1, 0), (1, 1)]
    want = resize_crop_bicubic(PO.prep_depth(depth, dmask, 1)[:, 0], px, step=sw.f)          # [S, h, w]
    dl = sw.depth_lat.cpu().numpy()[:, 0]
    for j, (i, _) in enumerate(items):
        assert np.abs(dl[j] - want[i]).max() < 1e-5


@pytest.mark.parametrize("Hs,Ws,res", [(96, 128, 64), (80, 80, 64), (50, 70, 64)])
def test_dataset_transforms_vs_the_reference_functions(Hs, Ws, res):
    """sweep.dataset_transforms = the per-sample tensors of HDF5Dataset.__getitem__ (E/dataset/dataset.py:229-271): the reference's
    apply_transforms_rgb / _mask are `Compose([Resize(res, BICUBIC), CenterCrop(res), Normalize([0.5], [0.5])])` on tensor / 255
    (:70-96) — evaluated here with torchvision exactly as written there — and apply_transforms_depth as restated in oracle/
    (pinned to the reference's own function by prep_golden.npz) followed by the same resize."""
    tv = pytest.importorskip("torchvision")
    from torchvision import transforms
    from mirrorfusion_b200.sweep import dataset_transforms
    from oracle import prep_oracle as PO
    from oracle.resize_oracle import resize_crop_bicubic
    rng = np.random.default_rng(Hs + Ws)
    N = 2
    rgb = rng.integers(0, 256, (N, Hs, Ws, 3), dtype=np.uint8)
    mask = np.zeros((N, Hs, Ws), np.uint8)
    mask[:, 10:40, 12:50] = 255
    masked = rgb.copy()
    masked[mask == 255] = 0                                                   # get_masked_image (:60-67)
    depth = (rng.random((N, Hs, Ws), dtype=np.float32) * 4 + 0.5).astype(np.float32)
    ops.lib()
    got = dataset_transforms(torch.from_numpy(rgb).cuda(), torch.from_numpy(masked).cuda(), torch.from_numpy(mask).cuda(),
                             torch.from_numpy(depth).cuda(), resolution=res)
    t_rgb = transforms.Compose([transforms.Resize(res, interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(res),
                                transforms.Normalize([0.5], [0.5])])
    t_mask = transforms.Compose([transforms.Resize(res, interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(res)])
    for i in range(N):
        want = t_rgb(torch.tensor(rgb[i], dtype=torch.float32).permute(2, 0, 1) / 255.0)
        assert (got["pixel_values"][i].cpu() - want).abs().max().item() < 2e-5
        want = t_rgb(torch.tensor(masked[i], dtype=torch.float32).permute(2, 0, 1) / 255.0)
        assert (got["conditioning_pixel_values"][i].cpu() - want).abs().max().item() < 2e-5
        want = t_mask((torch.tensor(mask[i], dtype=torch.float32) / 255.0).unsqueeze(0))
        assert (got["masks"][i].cpu() - want).abs().max().item() < 2e-5
    want_d = resize_crop_bicubic(PO.prep_depth(depth, mask, 1)[:, 0], res)
    assert np.abs(got["depth"].cpu().numpy()[:, 0] - want_d).max() < 2e-5
