"""Pins oracle/train_oracle.py (the CPU restatement of the fine-tune step's glue): against vectors produced by the reference's
OWN DDPMScheduler.add_noise / get_velocity / compute_snr and loss block (tests/golden/train_glue.npz, oracle/make_golden_train.py),
and against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW themselves (the calls E/train_brushnet_mirror.py:1460-1464 makes).
Also the host-only weight transform of the conv data gradient."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import train_oracle as T


def _gold(golden_dir):
    return np.load(os.path.join(golden_dir, "train_glue.npz"))


def test_noise_schedule_and_add_noise_match_reference(golden_dir):
    g = _gold(golden_dir)
    acp = T.alphas_cumprod()
    assert np.array_equal(acp, g["alphas_cumprod"])
    assert np.array_equal(T.add_noise(g["x0"], g["noise"], g["t"], acp), g["noisy"])
    assert np.array_equal(T.get_velocity(g["x0"], g["noise"], g["t"], acp), g["velocity"])
    np.testing.assert_allclose(T.compute_snr(g["t"], acp), g["snr"], rtol=1e-6)


def test_loss_and_gradient_match_reference_autograd(golden_dir):
    g = _gold(golden_dir)
    acp = T.alphas_cumprod()
    for name, w in (("plain", None), ("snr5", T.snr_weights(g["t"], acp, 5.0))):
        if w is not None:
            np.testing.assert_allclose(w, g["w_snr5"], rtol=1e-6)
        loss, per, grad = T.mse_loss(g["pred"], g["noise"], w)
        assert abs(loss - float(g[f"loss_{name}"])) < 2e-6 * abs(loss)
        np.testing.assert_allclose(grad, g[f"grad_{name}"], rtol=2e-5, atol=1e-9)


def test_clip_and_adamw_match_torch():
    gen = torch.Generator().manual_seed(5)
    shapes = [(7, 5), (33,), (4, 3, 3, 3)]
    params = [torch.nn.Parameter(torch.randn(s, generator=gen)) for s in shapes]
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    p = [q.detach().numpy().copy() for q in params]
    m = [np.zeros(s) for s in shapes]
    v = [np.zeros(s) for s in shapes]
    for step in range(1, 6):
        grads = [torch.randn(s, generator=gen) * (3.0 if step % 2 else 0.01) for s in shapes]
        for q, gr in zip(params, grads):
            q.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        coef, tot = T.clip_coef([gr.numpy() for gr in grads], 1.0)
        assert abs(tot - float(total)) < 1e-5 * tot
        for i in range(len(shapes)):
            p[i], m[i], v[i] = T.adamw_step(p[i], grads[i].numpy() * coef, m[i], v[i], step=step, lr=1e-3)
            np.testing.assert_allclose(p[i], params[i].detach().numpy(), rtol=3e-6, atol=1e-7)


def test_dgrad_weight_transform_is_the_data_gradient():
    # the packed dgrad weight, unpacked back to OIHW and applied as a FORWARD conv to dy, equals autograd's dx
    import mirrorfusion_b200.ops as ops
    gen = torch.Generator().manual_seed(9)
    for k in (3, 1):
        x = torch.randn(2, 16, 7, 9, generator=gen)
        w = torch.randn(24, 16, k, k, generator=gen)
        dy = torch.randn(2, 24, 7, 9, generator=gen)
        dx, dw, db = T.conv_grads(x.numpy(), w.numpy(), dy.numpy())
        with ops.precision("fp32"):
            wp = ops.pack_conv_dgrad_weight(w)                        # [Cin_fwd, k*k*Cout_fwd], K order (kh, kw, c)
        assert tuple(wp.shape) == (16, k * k * 24)
        w_back = wp.view(16, k, k, 24).permute(0, 3, 1, 2)
        got = F.conv2d(dy, w_back, padding=k // 2)
        np.testing.assert_allclose(got.numpy(), dx, rtol=1e-4, atol=1e-4)
        # and the wgrad K order the kernel writes: dw[co][(kh, kw, ci)]
        assert dw.transpose(0, 2, 3, 1).reshape(24, -1).shape == (24, k * k * 16)
