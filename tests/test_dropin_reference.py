"""The shipped drop-in (mirrorfusion_b200/dropin.py) against the REFERENCE's own classes — runs where /root/reference is mounted
(this container; the GPU box has no reference, there tests/test_gpu_dropin.py exercises the same forward bodies on a stand-in base).

What a maintainer relies on:
  * the re-classed modules pass the pipeline's hard `isinstance(self.brushnet, BrushNetModel)` gates (pipeline_brushnet.py:634-641,649,1115);
  * `prepare_extra_step_kwargs` picks the B200 schedulers' `step` kwargs by signature inspection (:556-571);
  * `Attention.forward` filters cross_attention_kwargs by the PROCESSOR's signature (attention_processor.py:517-531) and reaches ours;
  * nothing silently falls back: without a B200 the first forward raises MfbError.
"""
import os
import sys

import pytest
import torch

REF_SRC = "/root/reference/MirrorFusion/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    import transformers.utils as tu
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):          # removed in transformers 5, imported by pipeline_loading_utils.py:44
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    sys.path.insert(0, REF_SRC)
    try:
        import diffusers
        yield diffusers
    finally:
        sys.path.remove(REF_SRC)
        for k in [k for k in sys.modules if k == "diffusers" or k.startswith("diffusers.")]:
            del sys.modules[k]


def _reference_pipeline(diffusers):
    from mirrorfusion_b200.config import MICRO as cfg
    down = tuple("CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg.down_has_attn)
    up = tuple("CrossAttnUpBlock2D" if a else "UpBlock2D" for a in cfg.up_has_attn)
    torch.manual_seed(0)
    unet = diffusers.UNet2DConditionModel(sample_size=cfg.sample_size, in_channels=4, out_channels=4, block_out_channels=cfg.block_out_channels,
                                          layers_per_block=cfg.layers_per_block, down_block_types=down, up_block_types=up,
                                          cross_attention_dim=cfg.cross_attention_dim, attention_head_dim=cfg.heads,
                                          norm_num_groups=cfg.norm_num_groups).eval()
    bn = diffusers.BrushNetModel.from_unet(unet, conditioning_channels=cfg.conditioning_channels).eval()
    vae = diffusers.AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 2,
                                  up_block_types=("UpDecoderBlock2D",) * 2, block_out_channels=(32, 64), layers_per_block=1,
                                  latent_channels=4, norm_num_groups=8, sample_size=64).eval()
    base = diffusers.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                                   set_alpha_to_one=False, steps_offset=1)
    sched = diffusers.UniPCMultistepScheduler.from_config(base.config)                  # E/test_brushnet.py:158
    pipe = diffusers.StableDiffusionBrushNetPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, brushnet=bn, scheduler=sched,
                                                     safety_checker=None, feature_extractor=None, requires_safety_checker=False)
    return pipe, cfg


def test_install_passes_the_reference_pipelines_isinstance_gates(ref):
    from mirrorfusion_b200 import dropin
    from mirrorfusion_b200._lib import MfbError
    from mirrorfusion_b200.schedulers import B200UniPCScheduler
    pipe, cfg = _reference_pipeline(ref)
    keys = (set(pipe.unet.state_dict()), set(pipe.brushnet.state_dict()))
    dropin.install(pipe, ref)
    assert isinstance(pipe.brushnet, ref.BrushNetModel) and isinstance(pipe.unet, ref.UNet2DConditionModel)
    assert type(pipe.brushnet).__name__ == "BrushNetB200" and type(pipe.unet).__name__ == "UNetB200"
    assert (set(pipe.unet.state_dict()), set(pipe.brushnet.state_dict())) == keys          # still the reference modules' parameters
    import dataclasses
    assert dropin.net_config_of(pipe.unet, "unet") == cfg
    assert dataclasses.replace(dropin.net_config_of(pipe.brushnet, "brushnet"), sample_size=cfg.sample_size) == cfg   # BrushNetModel has no sample_size
    assert isinstance(pipe.scheduler, B200UniPCScheduler)
    # check_inputs: `assert False` unless brushnet IS-A BrushNetModel (:634-641); float scale gate (:649-650); image/mask/depth checks
    pe = torch.zeros(1, 77, cfg.cross_attention_dim)
    img = torch.zeros(1, 3, 64, 64)
    args = lambda scale: (None, img, img, None, None, pe, pe, None, None, scale, 0.0, 1.0, ["latents"])      # the call at :1045-1061
    pipe.check_inputs(*args(1.0), depth=img, normals=None)
    with pytest.raises(TypeError):
        pipe.check_inputs(*args(1), depth=img, normals=None)
    # prepare_extra_step_kwargs inspects scheduler.step's signature (:556-571): UniPC takes neither eta nor generator, DDIM both
    assert pipe.prepare_extra_step_kwargs(torch.Generator(), 0.0) == {}
    pipe.scheduler = dropin.b200_scheduler_for(ref.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                                                 clip_sample=False, set_alpha_to_one=False, steps_offset=1))
    with pytest.raises(NotImplementedError):                  # a config outside the path (DDIM's default clip_sample=True) is refused, loudly
        dropin.b200_scheduler_for(ref.DDIMScheduler())
    assert set(pipe.prepare_extra_step_kwargs(torch.Generator(), 0.0)) == {"eta", "generator"}
    # the pipeline's loop surface of the scheduler (:1171-1176,1257,1315)
    pipe.scheduler.set_timesteps(4, device="cpu")
    assert len(pipe.scheduler.timesteps) == 4 and pipe.scheduler.order == 1 and float(pipe.scheduler.init_noise_sigma) == 1.0
    # no silent fallback: this container has no B200, so the forward must fail loudly inside OUR code (not run the reference's)
    x = torch.zeros(2, 4, 8, 8)
    with pytest.raises(MfbError):
        pipe.brushnet(x, torch.tensor(10), encoder_hidden_states=torch.zeros(2, 77, cfg.cross_attention_dim),
                      brushnet_cond=torch.zeros(2, 6, 8, 8), conditioning_scale=1.0, guess_mode=False, return_dict=False)
    with pytest.raises(MfbError):
        pipe.unet(x, torch.tensor(10), encoder_hidden_states=torch.zeros(2, 77, cfg.cross_attention_dim), timestep_cond=None,
                  cross_attention_kwargs=None, down_block_add_samples=None, mid_block_add_sample=None, up_block_add_samples=None,
                  added_cond_kwargs=None, return_dict=False)


def test_attention_processor_on_the_reference_attention_class(ref):
    """Attention.forward drops cross_attention_kwargs the processor's __call__ does not name and passes the rest
    (attention_processor.py:517-531); set_processor / attn_processors / set_attn_processor accept ours (:380-398, unet :693-748)."""
    from diffusers.models.attention_processor import Attention
    from mirrorfusion_b200 import dropin
    from mirrorfusion_b200._lib import MfbError
    from mirrorfusion_b200.pipeline import B200AttnProcessor
    attn = Attention(query_dim=64, cross_attention_dim=32, heads=2, dim_head=32)
    proc = B200AttnProcessor()
    attn.set_processor(proc)
    assert attn.processor is proc
    hs, ehs = torch.zeros(1, 16, 64), torch.zeros(1, 77, 32)
    with pytest.raises(MfbError):                           # reached our kernels' front end: unknown kwarg filtered, `scale` / `temb` kept
        attn(hs, encoder_hidden_states=ehs, not_a_processor_kwarg=1, scale=1.0, temb=None)
    with pytest.raises(NotImplementedError):                # the mask argument arrives by name
        attn(hs, encoder_hidden_states=ehs, attention_mask=torch.zeros(1, 1, 77))
    pipe, _ = _reference_pipeline(ref)
    n = len(pipe.unet.attn_processors)
    p2 = dropin.install_attention_processor(pipe.unet)
    assert n == 32 and all(v is p2 for v in pipe.unet.attn_processors.values())     # 16 transformer blocks x (attn1, attn2)
