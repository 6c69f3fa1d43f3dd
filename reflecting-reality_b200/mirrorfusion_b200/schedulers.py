"""DDIM / UniPC schedulers of the MirrorFusion pipeline, re-expressed for one fused device kernel.

The reference computes each step with ~40 (UniPC) or ~12 (DDIM) tiny tensor ops driven by 0-dim CPU scalars
(S/schedulers/scheduling_unipc_multistep.py:385-833, S/schedulers/scheduling_ddim.py:344-466).  Every one of those
updates is a LINEAR combination of at most five tensors (x, last_sample, the two stored x0-predictions and the
new prediction), so the host only has to produce 12 scalars per step (`coefficients(i)`, computed in float64 from
the same sigma / alpha tables) and one elementwise kernel (`mfb_cfg_sched_step`) applies CFG, convert_model_output,
UniC and UniP in a single pass.

The classes keep the reference's scheduler surface — `set_timesteps`, `timesteps`, `init_noise_sigma`, `order`,
`scale_model_input`, `step(model_output, timestep, sample, return_dict=...)`, `config` — so they can be dropped
into the pipeline (which picks extra `step` kwargs by signature inspection, pipeline_brushnet.py:556-571).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from . import ops

# layout of the coefficient vector consumed by mfb_cfg_sched_step
G, C_X, C_EPS, A_LAST, A_M0, A_M1, A_MT, USE_CORR, B_X, B_MT, B_M0, B_EPS = range(12)


def _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule) -> np.ndarray:
    # float32 like the reference (torch.linspace(..., dtype=float32) ** 2, cumprod in float32)
    if beta_schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    elif beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    else:
        raise NotImplementedError(f"{beta_schedule} is not implemented")
    return torch.cumprod(1.0 - betas, dim=0).numpy()


class _Config(dict):
    """Scheduler config with key AND attribute access, like the reference's FrozenDict (S/configuration_utils.py:52-87): the
    reference's own `SomeScheduler.from_config(ours.config)` (E/test_brushnet.py:158) takes it as the dict it expects."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


class _Base:
    order = 1
    init_noise_sigma = 1.0

    def scale_model_input(self, sample, *args, **kwargs):
        """Identity for both schedulers (scheduling_ddim.py:238-253, scheduling_unipc_multistep.py:835-849)."""
        return sample

    # ---- device state for the fused kernel
    def _ensure_state(self, sample: torch.Tensor):
        shape = tuple(sample.shape)
        if getattr(self, "_state_shape", None) != shape or self._x.device != sample.device:
            z = lambda: torch.zeros(shape, device=sample.device, dtype=torch.float32)
            self._x, self._last, self._m0, self._m1 = z(), z(), z(), z()
            self._coef = torch.zeros(12, device=sample.device, dtype=torch.float32)
            self._state_shape = shape

    def coefficient_table(self, guidance_scale: float, eta: float = 0.0) -> torch.Tensor:
        rows = []
        for i in range(self.num_inference_steps):
            c = self.coefficients(i, eta) if eta else self.coefficients(i)
            c[G] = guidance_scale
            rows.append(c)
        return torch.tensor(np.stack(rows), dtype=torch.float32)

    # eta > 0 (stochastic DDIM): the variance noise rides in the kernel's `m0` operand, which DDIM does not use otherwise
    takes_variance_noise = False

    @staticmethod
    def variance_noise(shape, generator, device, dtype=torch.float32) -> torch.Tensor:
        """The draw of `randn_tensor(model_output.shape, generator=generator, device=..., dtype=...)` (scheduling_ddim.py:455-458,
        S/utils/torch_utils.py:39-88): a CPU generator draws on the host and the tensor is moved, a device generator (or none) draws
        on the device."""
        if isinstance(generator, (list, tuple)):
            if len(generator) != shape[0]:
                raise ValueError(f"got {len(generator)} generators for a batch of {shape[0]}")
            return torch.cat([_Base.variance_noise((1,) + tuple(shape[1:]), g, device, dtype) for g in generator], 0)
        if generator is not None and generator.device.type == "cpu":
            return torch.randn(tuple(shape), generator=generator, dtype=dtype).to(device)
        return torch.randn(tuple(shape), generator=generator, device=device, dtype=dtype)

    def _step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0, return_dict: bool = True,
              generator=None, variance_noise: Optional[torch.Tensor] = None):
        """Reference-compatible single step on an already guided `model_output` (same kernel, g = 0)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if eta != 0.0 and not self.takes_variance_noise:
            raise NotImplementedError("eta is a DDIM parameter (the reference ignores it for other schedulers, pipeline_brushnet.py:556-571)")
        if not sample.is_cuda:
            raise RuntimeError("mirrorfusion_b200 schedulers run on the GPU only (no CPU fallback)")
        i = self._index_for(timestep)
        self._ensure_state(sample)
        self._x.copy_(sample)
        eps = model_output.to(torch.float32).contiguous()
        c = self.coefficients(i, eta) if eta else self.coefficients(i)
        c[G] = 0.0
        if eta:
            if variance_noise is None:
                variance_noise = self.variance_noise(model_output.shape, generator, model_output.device, model_output.dtype)
            self._m0.copy_(variance_noise.to(torch.float32))
        self._coef.copy_(torch.tensor(c, dtype=torch.float32), non_blocking=False)
        ops.cfg_sched_step(eps, eps, self._x, self._last, self._m0, self._m1, self._coef)
        self._advance(i)
        prev = self._x.to(sample.dtype).clone()
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)


class B200DDIMScheduler(_Base):
    """DDIMScheduler (S/schedulers/scheduling_ddim.py): leading spacing + steps_offset (:299-342), the step (:404-466) incl. eta > 0:
    sigma_t = eta * sqrt((1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)) (:260-267,426-428), direction sqrt(1 - a_prev - sigma_t^2) eps
    (:444), + sigma_t * noise (:452-464)."""
    takes_variance_noise = True

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon"):
        if clip_sample or prediction_type != "epsilon":
            raise NotImplementedError("clip_sample / non-epsilon prediction are not on the MirrorFusion path")
        self.config = _Config(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type=prediction_type)
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else float(self.alphas_cumprod[0])
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config):
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "clip_sample", "set_alpha_to_one",
                "steps_offset", "prediction_type")
        get = (lambda k: config[k]) if isinstance(config, dict) else (lambda k: getattr(config, k))
        kw = {}
        for k in keys:
            try:
                kw[k] = get(k)
            except (KeyError, AttributeError):
                pass
        return cls(**kw)

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than {n_train}")
        self.num_inference_steps = num_inference_steps
        ratio = n_train // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self._ts = ts
        self.timesteps = torch.from_numpy(ts).to(device)

    def _index_for(self, timestep) -> int:
        t = int(timestep)
        idx = np.nonzero(self._ts == t)[0]
        if len(idx) == 0:
            raise ValueError(f"timestep {t} is not in the schedule")
        return int(idx[0])

    def _advance(self, i):
        pass

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False, generator=None,
             variance_noise=None, return_dict: bool = True):
        """DDIMScheduler.step (scheduling_ddim.py:344-466), same signature (the pipeline picks `eta` / `generator` by
        signature inspection, pipeline_brushnet.py:556-571)."""
        if use_clipped_model_output:
            raise NotImplementedError("use_clipped_model_output is not on the MirrorFusion path")
        if generator is not None and variance_noise is not None:
            raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                             " `variance_noise` stays `None`.")                              # :452-456
        return self._step(model_output, timestep, sample, eta=eta, return_dict=return_dict, generator=generator,
                          variance_noise=variance_noise)

    def coefficients(self, i: int, eta: float = 0.0) -> np.ndarray:
        t = int(self._ts[i])
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else self.final_alpha_cumprod
        c = np.zeros(12, dtype=np.float64)
        c[C_X] = 1.0 / math.sqrt(a_t)                       # x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t)   (:420)
        c[C_EPS] = -math.sqrt(1.0 - a_t) / math.sqrt(a_t)
        var = (1.0 - a_prev) / (1.0 - a_t) * (1.0 - a_t / a_prev)                               # _get_variance (:260-267)
        std = eta * math.sqrt(max(var, 0.0))                                                    # :428
        c[B_MT] = math.sqrt(a_prev)                         # x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev-std^2) eps (+ std noise)   (:444-464)
        c[B_EPS] = math.sqrt(max(1.0 - a_prev - std * std, 0.0))
        c[B_M0] = std                                       # the variance noise is handed over in the m0 operand
        return c


class B200UniPCScheduler(_Base):
    """UniPCMultistepScheduler (S/schedulers/scheduling_unipc_multistep.py) as MirrorFusion uses it
    (E/test_brushnet.py:158): bh2, predict_x0, solver_order 2, lower_order_final, epsilon, linspace spacing."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 solver_order=2, prediction_type="epsilon", predict_x0=True, solver_type="bh2", lower_order_final=True,
                 timestep_spacing="linspace", steps_offset=0, **_ignored):
        if solver_order != 2 or prediction_type != "epsilon" or not predict_x0 or solver_type != "bh2" \
                or timestep_spacing != "linspace":
            raise NotImplementedError("only the MirrorFusion UniPC configuration (order 2, bh2, x0-prediction) is implemented")
        self.config = _Config(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, solver_order=solver_order,
                                      prediction_type=prediction_type, predict_x0=predict_x0, solver_type=solver_type,
                                      lower_order_final=lower_order_final, timestep_spacing=timestep_spacing,
                                      steps_offset=steps_offset)
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps)[::-1].copy())
        self._step_index = None

    @classmethod
    def from_config(cls, config):
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "solver_order", "prediction_type",
                "predict_x0", "solver_type", "lower_order_final", "timestep_spacing")
        get = (lambda k: config[k]) if isinstance(config, dict) else (lambda k: getattr(config, k))
        kw = {}
        for k in keys:
            try:
                kw[k] = get(k)
            except (KeyError, AttributeError):
                pass
        return cls(**kw)

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        ts = np.linspace(0, n_train - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)   # :240-246
        ac = torch.from_numpy(self.alphas_cumprod)           # same float32 torch arithmetic as the reference
        sig = np.array(((1 - ac) / ac) ** 0.5)
        sigmas = np.interp(ts, np.arange(0, len(sig)), sig)                                                     # :272
        last = float(((1 - ac[0]) / ac[0]) ** 0.5)
        self.sigmas = np.concatenate([sigmas, [last]]).astype(np.float32)                                       # :273-274
        self._ts = ts
        self.timesteps = torch.from_numpy(ts).to(device)
        self.num_inference_steps = len(ts)
        self._step_index = None
        self.lower_order_nums = 0
        # order used by the predictor at step i / by the corrector at step i (= predictor order of step i-1), :810-817
        orders: List[int] = []
        lo = 0
        for i in range(self.num_inference_steps):
            o = min(self.config.solver_order, self.num_inference_steps - i) if self.config.lower_order_final \
                else self.config.solver_order
            o = min(o, lo + 1)
            orders.append(o)
            lo = min(lo + 1, self.config.solver_order)
        self._orders = orders

    def _index_for(self, timestep) -> int:
        if self._step_index is None:                      # _init_step_index / index_for_timestep :721-752
            t = int(timestep)
            idx = np.nonzero(self._ts == t)[0]
            self._step_index = len(self._ts) - 1 if len(idx) == 0 else int(idx[1] if len(idx) > 1 else idx[0])
        return self._step_index

    def _advance(self, i):
        self._step_index = i + 1

    @property
    def step_index(self):
        return self._step_index

    def step(self, model_output, timestep, sample, return_dict: bool = True):
        """UniPCMultistepScheduler.step (scheduling_unipc_multistep.py:754-833), same signature."""
        return self._step(model_output, timestep, sample, return_dict=return_dict)

    def _lam(self, k):
        s = float(self.sigmas[k])
        alpha = 1.0 / math.sqrt(s * s + 1.0)
        sigma = s * alpha
        return alpha, sigma, math.log(alpha) - math.log(sigma)

    @staticmethod
    def _bh(hh, order, rks):
        """R, b of the UniPC linear system (:529-553 / :662-686) for bh2; returns (h_phi_1, B_h, R, b)."""
        h_phi_1 = math.expm1(hh)
        B_h = math.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1.0
        fact = 1.0
        R, b = [], []
        for j in range(1, order + 1):
            R.append([rk ** (j - 1) for rk in rks])
            b.append(h_phi_k * fact / B_h)
            fact *= j + 1
            h_phi_k = h_phi_k / hh - 1.0 / fact
        return h_phi_1, B_h, np.array(R, dtype=np.float64), np.array(b, dtype=np.float64)

    def coefficients(self, i: int) -> np.ndarray:
        c = np.zeros(12, dtype=np.float64)
        a_i, s_i, lam_i = self._lam(i)
        c[C_X] = 1.0 / a_i                                  # x0 = (x - sigma_t eps) / alpha_t   (:425)
        c[C_EPS] = -s_i / a_i
        if i > 0:                                           # UniC with sigmas[i] (t) and sigmas[i-1] (s0), :599-719
            order = self._orders[i - 1]
            a_s0, s_s0, lam_s0 = self._lam(i - 1)
            h = lam_i - lam_s0
            rks = [(self._lam(i - (k + 1))[2] - lam_s0) / h for k in range(1, order)] + [1.0]
            hp1, Bh, R, b = self._bh(-h, order, rks)
            rhos = np.array([0.5]) if order == 1 else np.linalg.solve(R, b)
            c[USE_CORR] = 1.0
            c[A_LAST] = s_i / s_s0
            c[A_MT] = -a_i * Bh * rhos[-1]
            c[A_M0] = -a_i * hp1 + a_i * Bh * rhos[-1]
            if order == 2:
                c[A_M1] = -a_i * Bh * rhos[0] / rks[0]
                c[A_M0] += a_i * Bh * rhos[0] / rks[0]
        order = self._orders[i]                             # UniP with sigmas[i+1] (t) and sigmas[i] (s0), :455-582
        a_n, s_n, lam_n = self._lam(i + 1)
        h = lam_n - lam_i
        hp1, Bh = math.expm1(-h), math.expm1(-h)
        c[B_X] = s_n / s_i
        c[B_MT] = -a_n * hp1
        if order == 2:
            rk = (self._lam(i - 1)[2] - lam_i) / h
            c[B_M0] = -a_n * Bh * 0.5 / rk                  # rhos_p = 0.5 (:560); D1 = (m_prev - m_t) / rk
            c[B_MT] += a_n * Bh * 0.5 / rk
        return c
