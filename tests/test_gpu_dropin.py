"""The forward bodies of the shipped per-module drop-in (mirrorfusion_b200/dropin.py) on the device, over a STAND-IN for the
reference base classes (the GPU box has no reference; tests/test_dropin_reference.py runs the same factory over the reference's
own BrushNetModel / UNet2DConditionModel / StableDiffusionBrushNetPipeline where it is mounted)."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from mirrorfusion_b200.config import TINY
from mirrorfusion_b200.synth import make_inputs, make_state_dict


class _RefStub(torch.nn.Module):
    """What the drop-in reads from a reference ModelMixin: state_dict(), parameters() (device / dtype), .config (a dict)."""

    def __init__(self, sd, config):
        super().__init__()
        self._sd, self.config = sd, config
        self.anchor = torch.nn.Parameter(torch.zeros(1, device="cuda", dtype=torch.float16))

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def forward(self, *a, **k):
        raise AssertionError("the reference forward must never run once the drop-in is installed")


def _stand_in():
    class BrushNetModel(_RefStub):
        pass

    class UNet2DConditionModel(_RefStub):
        pass

    return types.SimpleNamespace(BrushNetModel=BrushNetModel, UNet2DConditionModel=UNet2DConditionModel)


def _config(cfg, net):
    base = dict(in_channels=4, block_out_channels=list(cfg.block_out_channels), layers_per_block=cfg.layers_per_block,
                attention_head_dim=cfg.heads, cross_attention_dim=cfg.cross_attention_dim, norm_num_groups=cfg.norm_num_groups,
                norm_eps=cfg.norm_eps)
    if net == "unet":
        base.update(out_channels=4, sample_size=cfg.sample_size, mid_block_type="UNetMidBlock2DCrossAttn",
                    down_block_types=["CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg.down_has_attn],
                    up_block_types=["CrossAttnUpBlock2D" if a else "UpBlock2D" for a in cfg.up_has_attn])
    else:
        base.update(conditioning_channels=cfg.conditioning_channels, mid_block_type="MidBlock2D",
                    down_block_types=["DownBlock2D"] * 4, up_block_types=["UpBlock2D"] * 4)
    return base


def test_dropin_subclasses_run_the_b200_forward_bit_identically():
    from mirrorfusion_b200 import dropin, pipeline as P
    cfg = TINY
    ns = _stand_in()
    usd, bsd = make_state_dict(cfg, "unet"), make_state_dict(cfg, "brushnet")
    unet, bn = ns.UNet2DConditionModel(usd, _config(cfg, "unet")), ns.BrushNetModel(bsd, _config(cfg, "brushnet"))
    class UniPCMultistepScheduler:                      # the drop-in dispatches on the reference scheduler's class name + config
        config = dict(P.B200UniPCScheduler().config)

    pipe = types.SimpleNamespace(unet=unet, brushnet=bn, scheduler=UniPCMultistepScheduler())
    dropin.install(pipe, ns)
    assert isinstance(pipe.brushnet, ns.BrushNetModel) and isinstance(pipe.unet, ns.UNet2DConditionModel)
    assert isinstance(pipe.scheduler, P.B200UniPCScheduler)
    inp = make_inputs(cfg, 2, seed=4)
    x = torch.cat([inp["latents"]] * 2).cuda().half()
    ehs, cond = inp["prompt_embeds"].cuda().half(), inp["conditioning_latents"].cuda().half()
    t = torch.tensor(431)
    # the calls of the reference loop (pipeline_brushnet.py:1277-1307), keyword for keyword
    d, m, u = pipe.brushnet(x, t, encoder_hidden_states=ehs, brushnet_cond=cond, conditioning_scale=0.8, guess_mode=False, return_dict=False)
    assert len(d) == 12 and len(u) == 15 and d[0].dtype == x.dtype
    keep = [a.clone() for a in list(d) + [m] + list(u)]
    eps = pipe.unet(x, t, encoder_hidden_states=ehs, timestep_cond=None, cross_attention_kwargs=None, down_block_add_samples=d,
                    mid_block_add_sample=m, up_block_add_samples=u, added_cond_kwargs=None, return_dict=False)[0]
    assert len(d) == 0 and len(u) == 0                               # consumed with pop(0) like the reference
    # the same through the plain classes
    bn2, un2 = P.B200BrushNetModel(bsd, cfg, dtype=torch.float16), P.B200UNet2DConditionModel(usd, cfg, dtype=torch.float16)
    d2, m2, u2 = bn2(x, t, encoder_hidden_states=ehs, brushnet_cond=cond, conditioning_scale=0.8, return_dict=False)
    for a, b in zip(keep, list(d2) + [m2] + list(u2)):
        assert torch.equal(a, b)
    eps2 = un2(x, t, ehs, down_block_add_samples=d2, mid_block_add_sample=m2, up_block_add_samples=u2, return_dict=False)[0]
    assert torch.equal(eps, eps2) and torch.isfinite(eps).all()
    # rebind() picks up new weights
    bsd2 = {k: v * 1.01 for k, v in bsd.items()}
    pipe.brushnet._sd = bsd2
    pipe.brushnet.rebind()
    d3, _, _ = pipe.brushnet(x, t, encoder_hidden_states=ehs, brushnet_cond=cond, conditioning_scale=0.8, return_dict=False)
    assert not torch.equal(d3[3], keep[3])
