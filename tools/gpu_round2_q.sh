#!/bin/bash
# round 2q: dK/dV kernel with the per-column L / D prefetched one tile ahead
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_bf16.py -q -k "attention" 2>&1 | tail -3
timeout 300 python tools/bench_attn_bwd.py --batch 8 2>&1 | tee gpurun_out/r02q_attn_bwd_bench.txt | tail -8
