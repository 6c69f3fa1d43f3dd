#!/bin/bash
# One GPU call that refreshes the fine-tune (config 4) evidence: the pending kernels' first run, the net-level program in both
# precisions, the glue bench.  usage: tools/gpu_train_round.sh   (outputs under gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zz_train_net.py -q -m gpu --runxfail 2>&1 | tail -15 | tee gpurun_out/train_net_tests.log
timeout 200 python tools/check_brushnet_trainer.py --precision fp32 --config tiny 2>&1 | tail -12
timeout 200 python tools/check_brushnet_trainer.py --precision bf16 --config tiny 2>&1 | tail -12
timeout 300 python tools/check_brushnet_trainer.py --precision bf16 --config sd15 --no-check --batch 8 --size 64 2>&1 | tail -3
timeout 120 python tools/bench_train_glue.py 2>&1 | tail -14
