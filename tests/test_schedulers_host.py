"""Host logic of the fused scheduler: the 12 coefficients per step, applied with the kernel's arithmetic
(restated in numpy), must reproduce the reference scheduler trajectories (tests/golden/sched_traj.npz)."""
import os

import numpy as np
import pytest

from mirrorfusion_b200 import schedulers as S


def apply_kernel_math(c, eps_u, eps_c, x, last, m0, m1):
    """numpy restatement of cfg_sched_kernel (csrc/misc.cu) — test-only."""
    e = eps_u + c[S.G] * (eps_c - eps_u)
    mt = c[S.C_X] * x + c[S.C_EPS] * e
    xc = x
    if c[S.USE_CORR] != 0:
        xc = c[S.A_LAST] * last + c[S.A_M0] * m0 + c[S.A_M1] * m1 + c[S.A_MT] * mt
    xn = c[S.B_X] * xc + c[S.B_MT] * mt + c[S.B_M0] * m0 + c[S.B_EPS] * e
    return xn, xc, mt, m0


@pytest.mark.parametrize("kind,n", [("ddim", 4), ("ddim", 10), ("unipc", 5), ("unipc", 10), ("unipc", 50)])
def test_coefficients_reproduce_reference_trajectory(golden_dir, kind, n):
    g = np.load(os.path.join(golden_dir, "sched_traj.npz"))
    s = S.B200DDIMScheduler() if kind == "ddim" else S.B200UniPCScheduler()
    s.set_timesteps(n)
    assert np.array_equal(s.timesteps.numpy(), g[f"{kind}{n}_timesteps"])
    if kind == "unipc":
        assert np.array_equal(s.sigmas, g[f"{kind}{n}_sigmas"])
    x = g["x0"].astype(np.float32)
    last = np.zeros_like(x); m0 = np.zeros_like(x); m1 = np.zeros_like(x)
    table = s.coefficient_table(0.0).numpy()
    assert table.shape == (n, 12)
    for i, t in enumerate(s.timesteps.numpy()):
        eps = (np.sin(3.0 * x + 0.01 * float(t)) * 0.9 + 0.1 * x).astype(np.float32)
        x, last, m0, m1 = [a.astype(np.float32) for a in apply_kernel_math(table[i], eps, eps, x, last, m0, m1)]
        ref = g[f"{kind}{n}_traj"][i]
        err = np.linalg.norm(x - ref) / np.linalg.norm(ref)
        assert err < 5e-6, f"step {i}: {err}"


def test_cfg_combination_is_folded_into_the_coefficients():
    s = S.B200UniPCScheduler()
    s.set_timesteps(7)
    t = s.coefficient_table(7.5)
    assert np.allclose(t[:, S.G].numpy(), 7.5)
    assert t[0, S.USE_CORR] == 0 and bool((t[1:, S.USE_CORR] == 1).all())
    # order warm-up 1 -> 2 and lower_order_final on the last step (scheduling_unipc_multistep.py:810-817)
    assert s._orders == [1, 2, 2, 2, 2, 2, 1]
    assert t[0, S.B_M0] == 0 and t[-1, S.B_M0] == 0 and t[3, S.B_M0] != 0


def test_scheduler_surface_matches_reference_contract():
    for cls in (S.B200DDIMScheduler, S.B200UniPCScheduler):
        s = cls()
        assert s.init_noise_sigma == 1.0 and s.order == 1
        assert s.scale_model_input("x", 3) == "x"
        s2 = cls.from_config(s.config)
        assert s2.config.beta_start == s.config.beta_start
        with pytest.raises(ValueError):
            cls().step(None, 1, None)          # set_timesteps not called (same error class as the reference)
    # DDIM from a UniPC config and vice versa (E/test_brushnet.py:158 builds UniPC from the pipeline scheduler's config)
    u = S.B200UniPCScheduler.from_config(S.B200DDIMScheduler().config)
    assert u.config.beta_schedule == "scaled_linear"


@pytest.mark.parametrize("eta,n,seed", [(0.7, 6, 11), (1.0, 10, 5)])
def test_stochastic_ddim_coefficients_reproduce_reference_trajectory(golden_dir, eta, n, seed):
    """eta > 0 (scheduling_ddim.py:426-464): sigma_t in the B_M0 slot, the direction coefficient shrunk to sqrt(1 - a_prev - sigma^2),
    and the variance noise — drawn per step from the caller's generator exactly like `randn_tensor` does — handed over as `m0`."""
    import torch
    g = np.load(os.path.join(golden_dir, "sched_eta_traj.npz"))
    s = S.B200DDIMScheduler()
    s.set_timesteps(n)
    table = s.coefficient_table(0.0, eta).numpy()
    assert (table[:, S.B_M0] > 0).all() and (s.coefficient_table(0.0).numpy()[:, S.B_M0] == 0).all()
    gen = torch.Generator().manual_seed(seed)
    x = g["x0"].astype(np.float32)
    z = np.zeros_like(x)
    for i, t in enumerate(s.timesteps.numpy()):
        eps = (np.sin(3.0 * x + 0.01 * float(t)) * 0.9 + 0.1 * x).astype(np.float32)
        noise = s.variance_noise(x.shape, gen, "cpu").numpy()
        x = apply_kernel_math(table[i], eps, eps, x, z, noise, z)[0].astype(np.float32)
        ref = g[f"eta{eta}_n{n}_seed{seed}_traj"][i]
        err = np.linalg.norm(x - ref) / np.linalg.norm(ref)
        assert err < 5e-6, f"step {i}: {err}"
    # UniPC takes neither eta nor generator (the pipeline's signature inspection drops them, pipeline_brushnet.py:556-571)
    import inspect
    assert "eta" not in inspect.signature(S.B200UniPCScheduler.step).parameters
    assert {"eta", "generator"} <= set(inspect.signature(S.B200DDIMScheduler.step).parameters)
