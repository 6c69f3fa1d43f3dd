"""Per-module drop-in for the REFERENCE's own pipeline object.

`StableDiffusionBrushNetPipeline.check_inputs` hard-asserts `isinstance(self.brushnet, BrushNetModel)`
(S/pipelines/brushnet/pipeline_brushnet.py:634-641, again :649 and :1115), so a module-level swap has to be a real
subclass of the reference class.  The reference package is not a dependency of this one (and does not exist on the GPU
box), so the subclasses are made at run time from whatever `diffusers` module the caller has imported:

    import diffusers                                    # the MirrorFusion fork
    from mirrorfusion_b200.dropin import install
    install(pipe, diffusers)                            # pipe: a StableDiffusionBrushNetPipeline, already on the GPU

`install` re-classes `pipe.brushnet` / `pipe.unet` in place (parameters, config, hooks and device placement stay what
they were, `state_dict()` keeps working, `isinstance` checks pass), replaces the scheduler by its B200 counterpart built
`from_config` (E/test_brushnet.py:158) and leaves VAE / CLIP / image processor alone.  The forward bodies are
`pipeline.B200BrushNetModel.forward` / `B200UNet2DConditionModel.forward`: reference signatures, fresh tap lists that
the UNet consumes with `pop(0)` (S/models/unets/unet_2d_condition.py:1218,1228,1306), reference return conventions.
The engines are built lazily at the first forward from the module's CURRENT `state_dict()`; `rebind()` after loading new
weights.  There is no fallback: off a B200 the first forward raises `MfbError`.
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

from .config import NetConfig
from .pipeline import B200AttnProcessor, B200BrushNetModel, B200UNet2DConditionModel
from .schedulers import B200DDIMScheduler, B200UniPCScheduler

_CACHE: Dict[int, Tuple[type, type]] = {}


def net_config_of(module, net: str) -> NetConfig:
    """NetConfig from the `.config` of a reference UNet2DConditionModel / BrushNetModel (same rules as a config.json)."""
    from .checkpoint import config_from_json
    cfg = dict(module.config)
    cfg.setdefault("_class_name", "UNet2DConditionModel" if net == "unet" else "BrushNetModel")
    return config_from_json(cfg, net)


class _B200Bound:
    """Mixin holding the lazily built B200 model; the reference class comes second in the MRO."""

    _b200_net = ""           # "unet" | "brushnet"
    _b200_impl_cls: Any = None

    def _b200_impl(self):
        impl = self.__dict__.get("_b200")
        if impl is None:
            p = next(self.parameters())
            cfg = net_config_of(self, self._b200_net)
            impl = self._b200_impl_cls({k: v.detach() for k, v in self.state_dict().items()}, cfg, device=p.device, dtype=p.dtype)
            self.__dict__["_b200"] = impl
        return impl

    def rebind(self):
        """Drop the packed weights / engines (call after `load_state_dict` or an optimizer step on this module)."""
        self.__dict__.pop("_b200", None)
        return self

    def forward(self, *args, **kwargs):
        return self._b200_impl().forward(*args, **kwargs)


def make_dropin(diffusers) -> Tuple[type, type]:
    """-> (BrushNetB200, UNetB200): subclasses of `diffusers.BrushNetModel` / `diffusers.UNet2DConditionModel` whose forward runs
    on the sm_100a kernels.  `diffusers` is the imported reference package (or any object with those two attributes)."""
    key = id(diffusers)
    if key not in _CACHE:
        bn = type("BrushNetB200", (_B200Bound, diffusers.BrushNetModel), {"_b200_net": "brushnet", "_b200_impl_cls": B200BrushNetModel})
        un = type("UNetB200", (_B200Bound, diffusers.UNet2DConditionModel), {"_b200_net": "unet", "_b200_impl_cls": B200UNet2DConditionModel})
        _CACHE[key] = (bn, un)
    return _CACHE[key]


def convert(module, diffusers):
    """Re-class an existing reference module in place; returns it."""
    bn_cls, un_cls = make_dropin(diffusers)
    if isinstance(module, diffusers.BrushNetModel):
        module.__class__ = bn_cls
    elif isinstance(module, diffusers.UNet2DConditionModel):
        module.__class__ = un_cls
    else:
        raise TypeError(f"{type(module).__name__} is neither a BrushNetModel nor a UNet2DConditionModel")
    return module.rebind()


def b200_scheduler_for(scheduler):
    """The B200 scheduler with the reference scheduler's config (DDIM / UniPC: the two the path uses)."""
    name = type(scheduler).__name__
    if "UniPC" in name:
        return B200UniPCScheduler.from_config(scheduler.config)
    if "DDIM" in name:
        return B200DDIMScheduler.from_config(scheduler.config)
    raise NotImplementedError(f"{name}: only DDIMScheduler and UniPCMultistepScheduler are on the MirrorFusion path")


def install(pipe, diffusers, scheduler: bool = True):
    """Swap the two networks (and the scheduler) of a reference StableDiffusionBrushNetPipeline for the B200 drop-ins."""
    convert(pipe.brushnet, diffusers)
    convert(pipe.unet, diffusers)
    if scheduler:
        pipe.scheduler = b200_scheduler_for(pipe.scheduler)
    return pipe


def install_attention_processor(unet):
    """`unet.set_attn_processor(B200AttnProcessor())` (unet_2d_condition.py:716-748): keeps the reference UNet's own forward and
    only routes every Attention's q/k/v projections + SDPA + out projection through the kernels."""
    proc = B200AttnProcessor()
    unet.set_attn_processor(proc)
    return proc
