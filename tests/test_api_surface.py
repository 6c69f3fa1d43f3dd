"""Drop-in surface: our entry points keep the parameter names/order/defaults of the reference ones
(tests/golden/reference_signatures.json, written by oracle/make_golden.py from the reference itself)."""
import inspect
import json
import os

import pytest

from mirrorfusion_b200 import pipeline as P
from mirrorfusion_b200 import schedulers as S

PAIRS = {
    "BrushNetModel.forward": P.B200BrushNetModel.forward,
    "UNet2DConditionModel.forward": P.B200UNet2DConditionModel.forward,
    "StableDiffusionBrushNetPipeline.__call__": P.MirrorFusionB200Pipeline.__call__,
    "AttnProcessor2_0.__call__": P.B200AttnProcessor.__call__,
    "UniPCMultistepScheduler.step": S.B200UniPCScheduler.step,
    "DDIMScheduler.step": S.B200DDIMScheduler.step,
    "UniPCMultistepScheduler.set_timesteps": S.B200UniPCScheduler.set_timesteps,
    "UniPCMultistepScheduler.scale_model_input": S.B200UniPCScheduler.scale_model_input,
}


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "reference_signatures.json")))


@pytest.mark.parametrize("name", sorted(PAIRS))
def test_signature_matches_reference(golden, name):
    ref = golden[name]
    ours = list(inspect.signature(PAIRS[name]).parameters.values())
    assert [p.name for p in ours] == [p["name"] for p in ref], name
    for o, r in zip(ours, ref):
        assert str(o.kind) == r["kind"], (name, o.name)
        if r["default"] is not None and r["default"] in ("None", "True", "False", "1.0", "0.0", "50", "7.5", "1", "'pil'"):
            if r["default"] == "'pil'":       # we return latents unless a vae_decode callable is supplied
                continue
            assert repr(o.default) == r["default"], (name, o.name, o.default, r["default"])


def test_golden_signatures_are_current_when_reference_is_mounted(golden):
    ref_src = "/root/reference/MirrorFusion/src"
    if not os.path.isdir(ref_src):
        pytest.skip("reference not mounted")
    import sys
    import transformers.utils as tu
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    sys.path.insert(0, ref_src)
    try:
        import diffusers
        got = [p for p in inspect.signature(diffusers.BrushNetModel.forward).parameters]
        assert got == [p["name"] for p in golden["BrushNetModel.forward"]]
        got = [p for p in inspect.signature(diffusers.UNet2DConditionModel.forward).parameters]
        assert got == [p["name"] for p in golden["UNet2DConditionModel.forward"]]
    finally:
        sys.path.remove(ref_src)


def test_pipeline_error_behaviour_mirrors_check_inputs():
    # pipeline_brushnet.py:573-693 (the subset that applies without tokenizer / PIL inputs); no GPU needed: the errors
    # are raised before any engine is built
    import torch
    pipe = P.MirrorFusionB200Pipeline({}, {}, device="cpu")
    pe = torch.zeros(1, 77, 768)
    with pytest.raises(ValueError):
        pipe(prompt_embeds=None)
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pe, negative_prompt_embeds=torch.zeros(1, 76, 768))
    with pytest.raises(TypeError):
        pipe(prompt_embeds=pe, brushnet_conditioning_scale=1)                      # must be a float (:649-650)
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pe, control_guidance_start=0.8, control_guidance_end=0.2)
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pe, control_guidance_end=1.5)
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pe, callback_on_step_end_tensor_inputs=["latents", "prompt_embeds"])
    with pytest.raises(NotImplementedError):
        pipe(prompt="a mirror", prompt_embeds=pe)
    with pytest.raises(TypeError):
        pipe(prompt_embeds=pe, not_an_argument=1)
    with pytest.raises(ValueError, match="does not support custom"):                  # retrieve_timesteps (:112-121), same message
        pipe(prompt_embeds=pe, negative_prompt_embeds=pe, timesteps=[999, 500, 1])
    with pytest.raises(ValueError, match="empty_prompt_embeds"):                      # no silent zero tensor for the unconditional half
        pipe(prompt_embeds=pe)
