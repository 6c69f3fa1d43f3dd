"""CPU restatement (numpy, float64 where it matters) of the training-step glue of the BrushNet fine-tune step —
TEST INFRASTRUCTURE ONLY (only tests/, smoke() and bench.py's cpu_baseline leg may import oracle/).

Follows E/train_brushnet_mirror.py:1404-1466 (E/ = /root/reference/MirrorFusion/examples/brushnet/):
  * DDPMScheduler.add_noise / get_velocity        S/schedulers/scheduling_ddpm.py:501-546
  * compute_snr + min-SNR loss weights            S/training_utils.py:50-73, train_brushnet_mirror.py:1433-1450
  * F.mse_loss(pred.float(), target.float())      train_brushnet_mirror.py:1433
  * accelerator.clip_grad_norm_ -> torch.nn.utils.clip_grad_norm_   train_brushnet_mirror.py:1460-1463
  * torch.optim.AdamW (single-tensor update)      train_brushnet_mirror.py:1190-1200,1464
Parity PINNED: tests/test_oracle_train.py checks add_noise / get_velocity / compute_snr against vectors produced by the
reference's own DDPMScheduler and compute_snr (tests/golden/train_glue.npz, written by oracle/make_golden_train.py),
and the clip + AdamW restatement against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW themselves (the reference
calls exactly those; torch is present wherever the tests run)."""
from __future__ import annotations

import numpy as np


def alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012) -> np.ndarray:
    """scaled_linear betas in float32 as the reference builds them (scheduling_ddpm.py:198-205)."""
    import torch
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).numpy()


def add_noise(x0: np.ndarray, noise: np.ndarray, t: np.ndarray, acp: np.ndarray) -> np.ndarray:
    a = acp[t].astype(np.float32)
    sa = (a ** 0.5).reshape(-1, *([1] * (x0.ndim - 1)))
    so = ((1 - a) ** 0.5).reshape(-1, *([1] * (x0.ndim - 1)))
    return (sa * x0 + so * noise).astype(np.float32)                       # scheduling_ddpm.py:524


def get_velocity(x0: np.ndarray, noise: np.ndarray, t: np.ndarray, acp: np.ndarray) -> np.ndarray:
    a = acp[t].astype(np.float32)
    sa = (a ** 0.5).reshape(-1, *([1] * (x0.ndim - 1)))
    so = ((1 - a) ** 0.5).reshape(-1, *([1] * (x0.ndim - 1)))
    return (sa * noise - so * x0).astype(np.float32)                       # scheduling_ddpm.py:545


def compute_snr(t: np.ndarray, acp: np.ndarray) -> np.ndarray:
    alpha = acp.astype(np.float32) ** 0.5
    sigma = (1.0 - acp.astype(np.float32)) ** 0.5
    return ((alpha[t] / sigma[t]) ** 2).astype(np.float32)                 # training_utils.py:72


def snr_weights(t: np.ndarray, acp: np.ndarray, snr_gamma: float, prediction_type: str = "epsilon") -> np.ndarray:
    snr = compute_snr(t, acp)
    w = np.minimum(snr, np.float32(snr_gamma))                             # train_brushnet_mirror.py:1440-1442
    if prediction_type == "epsilon":
        return (w / snr).astype(np.float32)
    if prediction_type == "v_prediction":
        return (w / (snr + 1)).astype(np.float32)
    raise ValueError(prediction_type)


def mse_loss(pred: np.ndarray, target: np.ndarray, weights: np.ndarray | None = None):
    """-> (loss, per_sample [B], d loss / d pred).  weights=None is F.mse_loss(reduction="mean")."""
    B = pred.shape[0]
    d = pred.astype(np.float64) - target.astype(np.float64)
    n = d[0].size
    per = (d.reshape(B, -1) ** 2).mean(1)
    w = np.ones(B) if weights is None else weights.astype(np.float64)
    loss = float((per * w).mean())
    grad = 2.0 * d * w.reshape(-1, *([1] * (pred.ndim - 1))) / (B * n)
    return loss, per, grad


def clip_coef(grads: list[np.ndarray], max_norm: float):
    """torch.nn.utils.clip_grad_norm_: total = || (||g_i||) ||_2 ; coef = min(1, max_norm / (total + 1e-6))."""
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads)))
    return min(1.0, max_norm / (total + 1e-6)), total


def adamw_step(p, g, m, v, *, step: int, lr=5e-6, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """torch/optim/adamw.py `_single_tensor_adamw` (no amsgrad, not maximize), float64.  Returns new (p, m, v)."""
    p, g, m, v = (a.astype(np.float64) for a in (p, g, m, v))
    p = p * (1 - lr * weight_decay)
    m = m + (1 - beta1) * (g - m)
    v = v * beta2 + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def conv_grads(x_nchw, w_oihw, dy_nchw):
    """Autograd of F.conv2d(x, w, padding=k//2) — what the reference's backward runs.  -> (dx, dw, dbias), float64."""
    import torch
    x = torch.tensor(x_nchw, dtype=torch.float64, requires_grad=True)
    w = torch.tensor(w_oihw, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(w.shape[0], dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x, w, b, padding=w.shape[-1] // 2)
    y.backward(torch.tensor(dy_nchw, dtype=torch.float64))
    return x.grad.numpy(), w.grad.numpy(), b.grad.numpy()


def groupnorm_silu_grads(x_nchw, gamma, beta, dy_nchw, groups, eps, silu=True):
    """Autograd of F.silu(F.group_norm(x, groups, gamma, beta, eps)) (S/models/resnet.py:337-338,381,393), float64.
    -> (dx, dgamma, dbeta)."""
    import torch
    x = torch.tensor(x_nchw, dtype=torch.float64, requires_grad=True)
    g = torch.tensor(gamma, dtype=torch.float64, requires_grad=True)
    b = torch.tensor(beta, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.group_norm(x, groups, g, b, eps)
    if silu:
        y = torch.nn.functional.silu(y)
    y.backward(torch.tensor(dy_nchw, dtype=torch.float64))
    return x.grad.numpy(), g.grad.numpy(), b.grad.numpy()


def resnet_block_grads(sd, prefix, x_nchw, emb, d_out_nchw, groups=32, eps=1e-5):
    """Autograd (float64) through the oracle's ResnetBlock2D restatement (oracle/mf_oracle.py `resnet`, pinned to the reference's
    block by tests/golden/resnet_block_grad.npz).  -> dict: out, dx, d_rowbias (gradient at time_emb_proj's OUTPUT), and the
    parameter gradients by state_dict name (OIHW)."""
    import torch
    from . import mf_oracle as O
    p = {k: v.detach().double().clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix + ".")}
    x = torch.as_tensor(x_nchw).double().clone().requires_grad_(True)
    e = torch.as_tensor(emb).double()
    rb = torch.nn.functional.linear(torch.nn.functional.silu(e), p[f"{prefix}.time_emb_proj.weight"], p[f"{prefix}.time_emb_proj.bias"])
    rb.retain_grad()
    # same algebra as O.resnet with the row bias exposed (it is the block boundary of the kernel program)
    F = torch.nn.functional
    h = F.silu(F.group_norm(x, groups, p[f"{prefix}.norm1.weight"], p[f"{prefix}.norm1.bias"], eps))
    h = F.conv2d(h, p[f"{prefix}.conv1.weight"], p[f"{prefix}.conv1.bias"], padding=1) + rb[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, p[f"{prefix}.norm2.weight"], p[f"{prefix}.norm2.bias"], eps))
    h = F.conv2d(h, p[f"{prefix}.conv2.weight"], p[f"{prefix}.conv2.bias"], padding=1)
    sc = x
    if f"{prefix}.conv_shortcut.weight" in p:
        sc = F.conv2d(x, p[f"{prefix}.conv_shortcut.weight"], p[f"{prefix}.conv_shortcut.bias"])
    out = sc + h
    with torch.no_grad():
        ref = O.resnet({k: v.detach() for k, v in p.items()}, prefix, x.detach(), e, groups, eps)
    assert torch.allclose(out.detach(), ref, rtol=1e-12, atol=1e-12)
    out.backward(torch.as_tensor(d_out_nchw).double())
    res = {"out": out.detach(), "dx": x.grad, "d_rowbias": rb.grad, "rowbias": rb.detach()}
    res.update({k: v.grad for k, v in p.items()})
    return res


def resnet_block_case(cin, cout, seed=77):
    """Seeded tensors of one ResnetBlock2D case (weights in state_dict naming under the prefix "r")."""
    import torch
    g = torch.Generator().manual_seed(seed)
    B, H, W, T = 2, 8, 8, 96
    rn = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale
    sd = {"r.norm1.weight": 1 + 0.2 * rn(cin), "r.norm1.bias": 0.1 * rn(cin),
          "r.conv1.weight": rn(cout, cin, 3, 3, scale=(9 * cin) ** -0.5), "r.conv1.bias": 0.1 * rn(cout),
          "r.time_emb_proj.weight": rn(cout, T, scale=T ** -0.5), "r.time_emb_proj.bias": 0.1 * rn(cout),
          "r.norm2.weight": 1 + 0.2 * rn(cout), "r.norm2.bias": 0.1 * rn(cout),
          "r.conv2.weight": rn(cout, cout, 3, 3, scale=(9 * cout) ** -0.5), "r.conv2.bias": 0.1 * rn(cout)}
    if cin != cout:
        sd["r.conv_shortcut.weight"] = rn(cout, cin, 1, 1, scale=cin ** -0.5)
        sd["r.conv_shortcut.bias"] = 0.1 * rn(cout)
    return sd, rn(B, cin, H, W), rn(B, T), rn(B, cout, H, W)


# ------------------------------------------------------------------------------------------------------------------------------
# Restatements of the three backward passes the frozen UNet's dgrad-only chain still needs in the FORM the
# kernels will compute them, so that the algorithms are pinned (tests/test_oracle_train.py, against torch autograd of the ops the
# reference calls) before any CUDA is written.

def attention_backward_two_pass(q, k, v, d_out, scale=None):
    """Backward of F.scaled_dot_product_attention (S/models/attention_processor.py:1266-1268; no mask, no dropout) as the two
    deterministic passes of a flash-style kernel pair, float64.  q [T, d], k / v [S, d] of ONE (batch, head).
      forward statistics per query i:  L_i = logsumexp_j(scale q_i.k_j),  O_i = sum_j p_ij v_j,  D_i = dO_i . O_i
      pass A (per query i):            p_ij = exp(scale q_i.k_j - L_i);  dS_ij = p_ij (dO_i.v_j - D_i);  dq_i = scale sum_j dS_ij k_j
      pass B (per key j):              dv_j = sum_i p_ij dO_i;            dk_j = scale sum_i dS_ij q_i
    Nothing of size T x S is stored; both passes recompute p from L.  -> (dq, dk, dv, L, D)."""
    q, k, v, d_out = (np.asarray(a, dtype=np.float64) for a in (q, k, v, d_out))
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    s = scale * q @ k.T
    m = s.max(1, keepdims=True)
    L = (m + np.log(np.exp(s - m).sum(1, keepdims=True)))[:, 0]
    dq = np.zeros_like(q)
    D = np.zeros(q.shape[0])
    for i in range(q.shape[0]):                              # pass A: one query per warp
        p = np.exp(scale * k @ q[i] - L[i])
        o = p @ v
        D[i] = d_out[i] @ o
        ds = p * (v @ d_out[i] - D[i])
        dq[i] = scale * ds @ k
    dk, dv = np.zeros_like(k), np.zeros_like(v)
    for j in range(k.shape[0]):                              # pass B: one key per warp, queries in order (deterministic)
        p = np.exp(scale * q @ k[j] - L)
        dv[j] = p @ d_out
        ds = p * (d_out @ v[j] - D)
        dk[j] = scale * ds @ q
    return dq, dk, dv, L, D


def layernorm_backward_dx(x, gamma, dy, eps=1e-5):
    """Data gradient of F.layer_norm over the last dim (S/models/attention.py:313,360,386; the UNet is frozen, so no dgamma / dbeta):
    with xhat = (x - mean) rstd and g = dy * gamma:  dx = rstd (g - mean(g) - xhat mean(g xhat)), row by row."""
    x, gamma, dy = (np.asarray(a, dtype=np.float64) for a in (x, gamma, dy))
    mean = x.mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(x.var(-1, keepdims=True) + eps)
    xhat = (x - mean) * rstd
    g = dy * gamma
    return rstd * (g - g.mean(-1, keepdims=True) - xhat * (g * xhat).mean(-1, keepdims=True))


def geglu_backward(proj, d_out):
    """Backward of GEGLU's `h, gate = proj.chunk(2, -1); h * gelu(gate)` (S/models/activations.py:100-103, exact erf GELU):
    d h = d out * gelu(gate);  d gate = d out * h * (Phi(gate) + gate phi(gate)).  -> d proj [.., 2*C]."""
    from math import erf, pi, sqrt
    proj, d_out = np.asarray(proj, dtype=np.float64), np.asarray(d_out, dtype=np.float64)
    C = proj.shape[-1] // 2
    h, gate = proj[..., :C], proj[..., C:]
    Phi = 0.5 * (1.0 + np.vectorize(erf)(gate / sqrt(2.0)))
    phi = np.exp(-0.5 * gate * gate) / sqrt(2.0 * pi)
    return np.concatenate([d_out * gate * Phi, d_out * h * (Phi + gate * phi)], -1)
