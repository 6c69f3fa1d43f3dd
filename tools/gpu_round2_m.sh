#!/bin/bash
# round 2m: eta > 0 / guess_mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MFB_PARITY_LOG=gpurun_out/r02m2_parity_metrics.jsonl
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropin.py tests/test_gpu_geometry.py -q 2>&1 | tail -40 > gpurun_out/r02m2_model_tests.log; tail -25 gpurun_out/r02m2_model_tests.log
