#!/bin/bash
# round 2l: vectorised bias-gradient column sums in the weight-gradient path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_zz_unet_train.py tests/test_gpu_zz_train_net.py -q 2>&1 | tail -8 > gpurun_out/r02l_train_tests.log; tail -4 gpurun_out/r02l_train_tests.log
timeout 900 python bench.py --train --steps 4 --warmup 3 > gpurun_out/r02l_train_b32.json 2> gpurun_out/r02l_train_b32.err; cut -c1-250 gpurun_out/r02l_train_b32.json; tail -3 gpurun_out/r02l_train_b32.err
