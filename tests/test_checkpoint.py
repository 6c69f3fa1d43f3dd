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
