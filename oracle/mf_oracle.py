"""CPU/fp32 ORACLE for the MirrorFusion denoising hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain-PyTorch functional restatement of what the reference computes for one
denoise step: BrushNetModel.forward -> UNet2DConditionModel.forward (with the 28 taps)
-> CFG combine -> scheduler step.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product path
(`mirrorfusion_b200`) never does, and fails loudly without its CUDA library.

Parity status: PINNED.  `oracle/make_golden.py` (run in the build container, where
/root/reference is mounted) loads the same seeded state_dicts into the reference's own
`UNet2DConditionModel` / `BrushNetModel` / schedulers and stores their outputs under
`tests/golden/`; `tests/test_oracle_golden.py` checks this file against those vectors
and against the known-answer values of the reference's own unit tests
(T/schedulers/test_scheduler_unipc.py:206-271, T/schedulers/test_scheduler_ddim.py:114-121,
T/models/test_layers_utils.py:94-120).

Path abbreviations in citations: S/ = /root/reference/MirrorFusion/src/diffusers/.
All tensors are NCHW like the reference; `sd` is a `state_dict()`-named dict of tensors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- embeddings
def timestep_embedding(timesteps: Tensor, dim: int) -> Tensor:
    """S/models/embeddings.py:27-67 with flip_sin_to_cos=True, downscale_freq_shift=0
    (Timesteps as configured by unet_2d_condition.py:286-290): fp32, [cos || sin]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def time_embed(sd: Dict[str, Tensor], t: Tensor, batch: int, c0: int, dtype) -> Tensor:
    """time_proj -> cast to model dtype -> TimestepEmbedding (embeddings.py:226-237;
    unet_2d_condition.py:1141-1155 / brushnet.py:750-772). `t` scalar or [B]."""
    t = torch.as_tensor(t)
    if t.dim() == 0:
        t = t[None]
    t = t.expand(batch)
    e = timestep_embedding(t, c0).to(dtype)
    e = F.linear(e, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    e = F.silu(e)
    return F.linear(e, sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])


# ----------------------------------------------------------------------------- layers
def resnet(sd, p: str, x: Tensor, emb: Tensor, groups: int, eps: float) -> Tensor:
    """ResnetBlock2D.forward, S/models/resnet.py:329-405 (default time_embedding_norm,
    dropout 0, output_scale_factor 1)."""
    h = F.silu(F.group_norm(x, groups, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps))
    h = F.conv2d(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    t = F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"])
    h = h + t[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps))
    h = F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def downsample(sd, p: str, x: Tensor) -> Tensor:
    """Downsample2D.forward, S/models/downsampling.py:134-154 (conv 3x3 stride 2 padding 1)."""
    return F.conv2d(x, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], stride=2, padding=1)


def upsample(sd, p: str, x: Tensor, size: Optional[Sequence[int]] = None) -> Tensor:
    """Upsample2D.forward, S/models/upsampling.py:145-186 (nearest x2 or to `size`, then conv 3x3)."""
    if size is None:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    else:
        x = F.interpolate(x, size=tuple(size), mode="nearest")
    return F.conv2d(x, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=1)


# The reference's DEFAULT processor is AttnProcessor2_0 = F.scaled_dot_product_attention (attention_processor.py:212-216,
# 1266-1268); AttnProcessor (:732-798: baddbmm + softmax + bmm) is the math it is equivalent to and what this restatement spells
# out.  The eager-GPU baseline (bench.py `gpu_eager_baseline`) flips this switch so that "the kernel to beat" runs the library
# flash / cuDNN attention the reference would run on a B200, not a materialised score matrix.
USE_SDPA = False


def attention(sd, p: str, x: Tensor, ctx: Optional[Tensor], heads: int) -> Tensor:
    """Attention + AttnProcessor2_0.__call__, S/models/attention_processor.py:1204-1286
    (no mask, q/k/v without bias, to_out[0] with bias, scale = dim_head^-0.5)."""
    src = x if ctx is None else ctx
    q = F.linear(x, sd[f"{p}.to_q.weight"])
    k = F.linear(src, sd[f"{p}.to_k.weight"])
    v = F.linear(src, sd[f"{p}.to_v.weight"])
    b, n, c = q.shape
    d = c // heads
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    else:
        s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
        o = torch.matmul(torch.softmax(s.float(), dim=-1).to(q.dtype), v)
    o = o.transpose(1, 2).reshape(b, n, c)
    return F.linear(o, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])


def transformer2d(sd, p: str, x: Tensor, ehs: Tensor, heads: int, groups: int) -> Tensor:
    """Transformer2DModel.forward continuous path (S/models/transformers/transformer_2d.py:334-346,
    378-430) with one BasicTransformerBlock (S/models/attention.py:291-412) and GEGLU FF
    (S/models/activations.py:100-103; exact erf GELU; first half value, second half gate)."""
    b, c, hh, ww = x.shape
    res = x
    h = F.group_norm(x, groups, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    h = F.conv2d(h, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    t = f"{p}.transformer_blocks.0"
    n = F.layer_norm(h, (c,), sd[f"{t}.norm1.weight"], sd[f"{t}.norm1.bias"], 1e-5)
    h = attention(sd, f"{t}.attn1", n, None, heads) + h
    n = F.layer_norm(h, (c,), sd[f"{t}.norm2.weight"], sd[f"{t}.norm2.bias"], 1e-5)
    h = attention(sd, f"{t}.attn2", n, ehs, heads) + h
    n = F.layer_norm(h, (c,), sd[f"{t}.norm3.weight"], sd[f"{t}.norm3.bias"], 1e-5)
    g = F.linear(n, sd[f"{t}.ff.net.0.proj.weight"], sd[f"{t}.ff.net.0.proj.bias"])
    val, gate = g.chunk(2, dim=-1)
    ff = F.linear(val * F.gelu(gate), sd[f"{t}.ff.net.2.weight"], sd[f"{t}.ff.net.2.bias"])
    h = ff + h
    h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
    h = F.conv2d(h, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    return h + res


# ----------------------------------------------------------------------------- BrushNet
def brushnet_forward(sd, cfg, sample: Tensor, t, brushnet_cond: Tensor, conditioning_scale: float = 1.0,
                     guess_mode: bool = False) -> Tuple[List[Tensor], Tensor, List[Tensor]]:
    """BrushNetModel.forward, S/models/brushnet.py:678-925 (MirrorFusion path: resnet-only blocks,
    no global pooling).  Returns (down taps[12], mid tap, up taps[15]) for SD1.5.
    guess_mode: tap k is scaled by conditioning_scale * logspace(-1, 0, 28)[k] (:896-902)."""
    G, eps = cfg.norm_num_groups, cfg.norm_eps
    boc = cfg.block_out_channels
    emb = time_embed(sd, t, sample.shape[0], boc[0], sample.dtype)
    x = torch.cat([sample, brushnet_cond], 1)                                   # brushnet.py:810
    x = F.conv2d(x, sd["conv_in_condition.weight"], sd["conv_in_condition.bias"], padding=1)
    down = [x]
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}", x, emb, G, eps)
            down.append(x)
        if i != len(boc) - 1:
            x = downsample(sd, f"down_blocks.{i}.downsamplers.0", x)
            down.append(x)
    down_taps = [F.conv2d(h, sd[f"brushnet_down_blocks.{k}.weight"], sd[f"brushnet_down_blocks.{k}.bias"])
                 for k, h in enumerate(down)]                                    # brushnet.py:831-834
    x = resnet(sd, "mid_block.resnets.0", x, emb, G, eps)                        # MidBlock2D unet_2d_blocks.py:1082-1111
    x = resnet(sd, "mid_block.resnets.1", x, emb, G, eps)
    mid_tap = F.conv2d(x, sd["brushnet_mid_block.weight"], sd["brushnet_mid_block.bias"])
    ups: List[Tensor] = []
    skips = list(down)
    for i in range(len(boc)):
        nl = cfg.layers_per_block + 1
        res = skips[-nl:]
        skips = skips[:-nl]
        for j in range(nl):
            x = torch.cat([x, res[-1 - j]], 1)                                    # UpBlock2D :2711-2728
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}", x, emb, G, eps)
            ups.append(x)
        if i != len(boc) - 1:
            x = upsample(sd, f"up_blocks.{i}.upsamplers.0", x, size=skips[-1].shape[2:])  # brushnet.py:864-865
            ups.append(x)
    up_taps = [F.conv2d(h, sd[f"brushnet_up_blocks.{k}.weight"], sd[f"brushnet_up_blocks.{k}.bias"])
               for k, h in enumerate(ups)]
    if guess_mode:                                                               # brushnet.py:896-902
        sc = torch.logspace(-1, 0, len(down_taps) + 1 + len(up_taps)) * conditioning_scale
        nd = len(down_taps)
        return [d * sc[k] for k, d in enumerate(down_taps)], mid_tap * sc[nd], [u * sc[nd + 1 + k] for k, u in enumerate(up_taps)]
    s = conditioning_scale                                                       # brushnet.py:904-906
    return [d * s for d in down_taps], mid_tap * s, [u * s for u in up_taps]


# ----------------------------------------------------------------------------- UNet
def unet_forward(sd, cfg, sample: Tensor, t, ehs: Tensor,
                 down_add: Optional[List[Tensor]] = None, mid_add: Optional[Tensor] = None,
                 up_add: Optional[List[Tensor]] = None) -> Tensor:
    """UNet2DConditionModel.forward, S/models/unets/unet_2d_condition.py:1039-1348, SD1.5 config, with the
    BrushNet tap sites (:1215-1218, 1288-1289 and unet_2d_blocks.py:1388-1398, 1483-1493, 2626-2635,
    2751-2761).  Unlike the reference, the tap lists are not mutated."""
    G, eps, H = cfg.norm_num_groups, cfg.norm_eps, cfg.heads
    boc = cfg.block_out_channels
    use = down_add is not None and mid_add is not None and up_add is not None
    da = list(down_add) if use else None
    ua = list(up_add) if use else None
    emb = time_embed(sd, t, sample.shape[0], boc[0], sample.dtype)
    x = F.conv2d(sample, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    skips = [x]                                                                   # pre-tap (:1215)
    if use:
        x = x + da.pop(0)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}", x, emb, G, eps)
            if cfg.down_has_attn[i]:
                x = transformer2d(sd, f"down_blocks.{i}.attentions.{j}", x, ehs, H, G)
            if use:
                x = x + da.pop(0)
            skips.append(x)
        if i != len(boc) - 1:
            x = downsample(sd, f"down_blocks.{i}.downsamplers.0", x)
            if use:
                x = x + da.pop(0)
            skips.append(x)
    x = resnet(sd, "mid_block.resnets.0", x, emb, G, eps)                         # UNetMidBlock2DCrossAttn :850-899
    x = transformer2d(sd, "mid_block.attentions.0", x, ehs, H, G)
    x = resnet(sd, "mid_block.resnets.1", x, emb, G, eps)
    if use:
        x = x + mid_add
    for i in range(len(boc)):
        nl = cfg.layers_per_block + 1
        res = skips[-nl:]
        skips = skips[:-nl]
        for j in range(nl):
            x = torch.cat([x, res[-1 - j]], 1)
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}", x, emb, G, eps)
            if cfg.up_has_attn[i]:
                x = transformer2d(sd, f"up_blocks.{i}.attentions.{j}", x, ehs, H, G)
            if use:
                x = x + ua.pop(0)
        if i != len(boc) - 1:
            # UNet passes upsample_size only for odd sizes (:1116-1126) -> scale_factor=2 path
            size = skips[-1].shape[2:] if (x.shape[2] * 2, x.shape[3] * 2) != tuple(skips[-1].shape[2:]) else None
            x = upsample(sd, f"up_blocks.{i}.upsamplers.0", x, size=size)
            if use:
                x = x + ua.pop(0)
    x = F.silu(F.group_norm(x, G, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], eps))
    return F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


def noise_pred_step(unet_sd, bn_sd, cfg, latent_in: Tensor, t, ehs: Tensor, cond: Tensor, scale: float = 1.0, guess_mode: bool = False):
    """Loop body S/pipelines/brushnet/pipeline_brushnet.py:1277-1307: BrushNet then UNet with taps.
    guess_mode (with CFG, :1262-1301): BrushNet sees the conditional half only (`cond` has half the batch), its taps are log-scaled,
    and the unconditional half of the UNet gets zeros."""
    if guess_mode:
        n = latent_in.shape[0] // 2
        d, m, u = brushnet_forward(bn_sd, cfg, latent_in[n:], t, cond[-n:], scale, guess_mode=True)
        z = lambda a: torch.cat([torch.zeros_like(a), a])
        d, m, u = [z(a) for a in d], z(m), [z(a) for a in u]
        return unet_forward(unet_sd, cfg, latent_in, t, ehs, d, m, u), (d, m, u)
    d, m, u = brushnet_forward(bn_sd, cfg, latent_in, t, cond, scale)
    return unet_forward(unet_sd, cfg, latent_in, t, ehs, d, m, u), (d, m, u)


def cfg_combine(noise_pred: Tensor, guidance_scale: float) -> Tensor:
    """pipeline_brushnet.py:1310-1312 — batch halves are [uncond, cond]."""
    u, c = noise_pred.chunk(2)
    return u + guidance_scale * (c - u)


# ----------------------------------------------------------------------------- schedulers
def _scaled_linear_alphas_cumprod(beta_start: float, beta_end: float, n: int) -> Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def _linear_alphas_cumprod(beta_start: float, beta_end: float, n: int) -> Tensor:
    betas = torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


class DDIMOracle:
    """DDIMScheduler (S/schedulers/scheduling_ddim.py): set_timesteps :299-342 (leading spacing),
    step :344-466 with eta=0, prediction_type epsilon."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, steps_offset=1):
        f = _scaled_linear_alphas_cumprod if beta_schedule == "scaled_linear" else _linear_alphas_cumprod
        self.alphas_cumprod = f(beta_start, beta_end, num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.n_train = num_train_timesteps
        self.clip_sample = clip_sample
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.order = 1

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.n_train // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, t=None):
        return sample

    def step(self, model_output: Tensor, t: int, sample: Tensor) -> Tensor:
        t = int(t)
        prev_t = t - self.n_train // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
        eps = model_output
        if self.clip_sample:
            x0 = x0.clamp(-1.0, 1.0)
            eps = (sample - a_t ** 0.5 * x0) / beta_t ** 0.5
        direction = (1 - a_prev) ** 0.5 * eps            # eta = 0 -> std_dev_t = 0
        return a_prev ** 0.5 * x0 + direction


class UniPCOracle:
    """UniPCMultistepScheduler (S/schedulers/scheduling_unipc_multistep.py), bh2, predict_x0, order 2,
    lower_order_final, epsilon prediction, linspace spacing: set_timesteps :229-293,
    convert_model_output :385-453, UniP :455-582, UniC :584-719, step :754-833.
    Scalars are 0-dim fp32 CPU tensors exactly like the reference (self.sigmas stays on CPU)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 solver_order=2, lower_order_final=True):
        f = _scaled_linear_alphas_cumprod if beta_schedule == "scaled_linear" else _linear_alphas_cumprod
        self.alphas_cumprod = f(beta_start, beta_end, num_train_timesteps)
        self.n_train = num_train_timesteps
        self.solver_order = solver_order
        self.lower_order_final = lower_order_final
        self.init_noise_sigma = 1.0
        self.order = 1

    def set_timesteps(self, n: int):
        ts = np.linspace(0, self.n_train - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sigmas = np.interp(ts, np.arange(0, len(sig)), sig)
        last = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.num_inference_steps = len(ts)
        self.model_outputs: List[Optional[Tensor]] = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = 1

    def scale_model_input(self, sample, t=None):
        return sample

    @staticmethod
    def _alpha_sigma(sigma):
        alpha = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha, sigma * alpha

    def _coeffs(self, s_t, s_s0, s_prev_list, order):
        """Common part of UniP/UniC: returns (sigma_t/sigma_s0, alpha_t, h_phi_1, B_h, rks, R, b)."""
        alpha_t, sigma_t = self._alpha_sigma(s_t)
        alpha_s0, sigma_s0 = self._alpha_sigma(s_s0)
        lam_t = torch.log(alpha_t) - torch.log(sigma_t)
        lam_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lam_t - lam_s0
        rks = []
        for s_i in s_prev_list:
            a_i, sg_i = self._alpha_sigma(s_i)
            rks.append((torch.log(a_i) - torch.log(sg_i) - lam_s0) / h)
        rks.append(1.0)
        rks_t = torch.tensor(rks)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)
        R, b = [], []
        fact = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks_t, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return sigma_t / sigma_s0, alpha_t, h_phi_1, B_h, rks, torch.stack(R), torch.tensor(b)

    def step(self, model_output: Tensor, t, sample: Tensor) -> Tensor:
        i = self.step_index
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i])
        m_t = (sample - sigma_t * model_output) / alpha_t                              # :425
        if i > 0 and self.last_sample is not None:                                   # UniC :584-719
            order = self.this_order
            m0 = self.model_outputs[-1]
            prev = [self.sigmas[i - (k + 1)] for k in range(1, order)]
            r, a_t, hp1, Bh, rks, R, b = self._coeffs(self.sigmas[i], self.sigmas[i - 1], prev, order)
            rhos_c = torch.tensor([0.5], dtype=sample.dtype) if order == 1 else torch.linalg.solve(R, b)
            x_ = r * self.last_sample - a_t * hp1 * m0
            corr = 0
            for k in range(1, order):
                corr = corr + rhos_c[k - 1] * ((self.model_outputs[-(k + 1)] - m0) / rks[k - 1])
            sample = (x_ - a_t * Bh * (corr + rhos_c[-1] * (m_t - m0))).to(sample.dtype)
        for k in range(self.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = m_t
        this_order = min(self.solver_order, len(self.timesteps) - i) if self.lower_order_final else self.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        order = self.this_order                                                      # UniP :455-582
        m0 = m_t
        prev = [self.sigmas[i - k] for k in range(1, order)]
        r, a_t, hp1, Bh, rks, R, b = self._coeffs(self.sigmas[i + 1], self.sigmas[i], prev, order)
        x_ = r * sample - a_t * hp1 * m0
        if order == 2:
            pred = 0.5 * ((self.model_outputs[-2] - m0) / rks[0])
        elif order == 1:
            pred = 0
        else:
            raise NotImplementedError("solver_order > 2 is not on the MirrorFusion path")
        out = (x_ - a_t * Bh * pred).to(sample.dtype)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return out


def denoise_loop(unet_sd, bn_sd, cfg, scheduler, latents: Tensor, ehs: Tensor, cond: Tensor, steps: int,
                 guidance_scale: float = 7.5, conditioning_scale: float = 1.0, return_trace: bool = False):
    """The hot loop, pipeline_brushnet.py:1249-1315 (do_classifier_free_guidance=True)."""
    scheduler.set_timesteps(steps)
    latents = latents * scheduler.init_noise_sigma
    trace = []
    for t in scheduler.timesteps:
        x_in = scheduler.scale_model_input(torch.cat([latents] * 2), t)
        eps, _ = noise_pred_step(unet_sd, bn_sd, cfg, x_in, t, ehs, cond, conditioning_scale)
        guided = cfg_combine(eps, guidance_scale)
        latents = scheduler.step(guided, t, latents)
        if return_trace:
            trace.append((eps, latents))
    return (latents, trace) if return_trace else latents
