#!/bin/bash
# round 2h: fp32 parity mode of the frozen-UNet chain / whole fine-tune step on the device
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MFB_PARITY_LOG=gpurun_out/r02h_parity_metrics.jsonl
rm -f $MFB_PARITY_LOG
timeout 1200 python -m pytest tests/test_gpu_zz_unet_train.py tests/test_gpu_zz_train_net.py tests/test_gpu_train_bf16.py tests/test_gpu_train.py -q 2>&1 | tail -60 > gpurun_out/r02h_train_tests.log; tail -30 gpurun_out/r02h_train_tests.log
