"""Checkpoint import (mirrorfusion_b200/checkpoint.py) against the on-disk layout the REFERENCE writes
(tests/golden/micro_checkpoint_layout.json: config.json + safetensors header of `save_pretrained`, produced by
oracle/make_golden.py, which also ran the loader on the reference-written files themselves)."""
import dataclasses
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from mirrorfusion_b200 import checkpoint as CK
from mirrorfusion_b200.config import MICRO, param_shapes
from mirrorfusion_b200.synth import make_state_dict


@pytest.fixture()
def layout(golden_dir):
    with open(os.path.join(golden_dir, "micro_checkpoint_layout.json")) as f:
        return json.load(f)


def _write(tmp_path, layout, net, sd, dtype=torch.float32):
    d = tmp_path / net
    d.mkdir()
    (d / "config.json").write_text(json.dumps(layout[net]["config"]))
    save_file({k: v.to(dtype).contiguous() for k, v in sd.items()}, str(d / CK.WEIGHTS[0]))
    return str(d)


def test_reference_layout_is_what_the_census_expects(layout):
    assert layout["loader_verified_on_reference_written_files"] is True
    for net in ("unet", "brushnet"):
        assert layout[net]["files"] == ["config.json", CK.WEIGHTS[0]]
        want = {k: list(s) for k, s in param_shapes(MICRO, net)}
        got = {k: v[1] for k, v in layout[net]["tensors"].items()}
        assert got == want                                   # names AND shapes of the reference's safetensors header
        assert dataclasses.asdict(CK.config_from_json(layout[net]["config"], net))["block_out_channels"] == MICRO.block_out_channels


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_round_trip_through_a_reference_shaped_directory(tmp_path, layout, dtype):
    usd, bsd = make_state_dict(MICRO, "unet"), make_state_dict(MICRO, "brushnet")
    cfg, lu, lb = CK.load_mirrorfusion(_write(tmp_path, layout, "unet", usd, dtype), _write(tmp_path, layout, "brushnet", bsd, dtype))
    assert dataclasses.asdict(cfg) == dataclasses.asdict(MICRO)
    assert all(v.dtype == torch.float32 for v in lu.values())
    assert all(torch.equal(lu[k], usd[k].to(dtype).float()) for k in usd)
    assert all(torch.equal(lb[k], bsd[k].to(dtype).float()) for k in bsd)


def test_errors_are_loud(tmp_path, layout):
    usd = make_state_dict(MICRO, "unet")
    bad = dict(usd)
    del bad["conv_in.bias"]
    bad["mid_block.resnets.0.conv1.weight"] = bad["mid_block.resnets.0.conv1.weight"][:, :, :1]
    bad["not.a.parameter"] = torch.zeros(1)
    with pytest.raises(CK.CheckpointError, match="missing .*conv_in.bias.*unexpected .*not.a.parameter.*wrong shape"):
        CK.load_model_dir(_write(tmp_path, layout, "unet", bad), "unet")
    with pytest.raises(CK.CheckpointError, match="expected 'BrushNetModel'"):
        CK.config_from_json(layout["unet"]["config"], "brushnet")
    lin = dict(layout["unet"]["config"], use_linear_projection=True)
    with pytest.raises(CK.CheckpointError, match="use_linear_projection"):
        CK.config_from_json(lin, "unet")
    with pytest.raises(CK.CheckpointError, match="no diffusion_pytorch_model"):
        (tmp_path / "empty").mkdir()
        ((tmp_path / "empty") / "config.json").write_text(json.dumps(layout["unet"]["config"]))
        CK.load_model_dir(str(tmp_path / "empty"), "unet")


def test_exported_brushnet_directory_loads_in_the_reference(tmp_path):
    """checkpoint.save_brushnet_dir writes what the reference's checkpoint hook writes (E/train_brushnet_mirror.py:997-1032): the
    reference's own `BrushNetModel.from_pretrained` rebuilds the module from our config.json and loads our safetensors strictly, and
    the config.json equals, key for key, the one `save_pretrained` of the reference wrote for the same architecture."""
    import json
    import sys
    from mirrorfusion_b200 import checkpoint as CK
    from mirrorfusion_b200.config import MICRO
    from mirrorfusion_b200.synth import make_state_dict
    sd = make_state_dict(MICRO, "brushnet", seed=3)
    d = str(tmp_path / "checkpoint-7" / "brushnet")
    CK.save_brushnet_dir(d, MICRO, sd)
    assert sorted(os.listdir(d)) == ["config.json", "diffusion_pytorch_model.safetensors"]
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "micro_checkpoint_layout.json")))
    assert json.load(open(os.path.join(d, "config.json"))) == gold["brushnet"]["config"]       # what the reference itself wrote
    cfg, back = CK.load_model_dir(d, "brushnet")
    assert all(torch.equal(back[k], sd[k]) for k in sd) and cfg.block_out_channels == MICRO.block_out_channels
    ref_src = "/root/reference/MirrorFusion/src"
    if not os.path.isdir(ref_src):
        return
    import transformers.utils as tu
    if not hasattr(tu, "FLAX_WEIGHTS_NAME"):
        tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
    saved = {k: v for k, v in sys.modules.items() if k == "diffusers" or k.startswith("diffusers.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, ref_src)
    try:
        import diffusers
        bn = diffusers.BrushNetModel.from_pretrained(d)              # the reference's own loader, strict
        rsd = bn.state_dict()
        assert set(rsd) == set(sd) and all(torch.equal(rsd[k], sd[k]) for k in sd)
    finally:
        sys.path.remove(ref_src)
        for k in [k for k in sys.modules if k == "diffusers" or k.startswith("diffusers.")]:
            del sys.modules[k]
        sys.modules.update(saved)
