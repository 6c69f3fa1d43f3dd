"""The baseline leg of bench.py: the UNMODIFIED reference (baseline/_ref or the /root/reference mount) behind its own
StableDiffusionBrushNetPipeline.__call__ (baseline/reference_arm.py).  Runs where a reference tree is reachable; TINY nets so that
it takes seconds."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_runs_the_reference_pipeline():
    from baseline import reference_arm as R
    path, kind = R.find_reference()
    if path is None:
        pytest.skip(kind)
    saved = {k: v for k, v in sys.modules.items() if k == "diffusers" or k.startswith("diffusers.")}
    try:
        from mirrorfusion_b200.config import TINY
        r = R.run(TINY, 1, 16, steps=2, warmup=1, device="cpu", dtype=torch.float32, threads=2)
        import diffusers
        assert os.path.abspath(diffusers.__file__).startswith(os.path.abspath(path))          # the reference's own package, not ours
        assert diffusers.StableDiffusionBrushNetPipeline.__module__.startswith("diffusers.pipelines.brushnet")
        assert r["steps_timed"] == 2 and r["sec_per_step"] > 0 and r["latents_finite"] and r["latents_shape"] == [1, 4, 16, 16]
    finally:
        for k in [k for k in sys.modules if k == "diffusers" or k.startswith("diffusers.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        if path in sys.path:
            sys.path.remove(path)


def test_bench_reference_line_contract(monkeypatch, capsys):
    """`bench.py --impl reference` prints ONE JSON line with the arm's keys (metric / unit / config of our arm, impl, cpu_baseline,
    e2e with zero copies) — checked here on a stubbed timing so that no SD1.5-sized net is built on the CPU suite."""
    import json
    import bench
    monkeypatch.setattr(bench, "cpu_reference_run", lambda steps, warmup, images=1: {
        "sec_per_step": 2.0, "images_per_s": images / (50 * 2.0), "cores": 4, "kind": "reference", "sample": "stub"})
    args = type("A", (), {"steps": 3, "warmup": 1, "latent": 64, "gpus": 1})()
    bench.run_reference(args, 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.metric_name(64) and line["unit"] == "images/s"
    assert line["steps"] == 3 and line["warmup"] == 1 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    bench.run_reference(args, 1)                                  # other ranks print nothing
    assert capsys.readouterr().out.strip() == ""
