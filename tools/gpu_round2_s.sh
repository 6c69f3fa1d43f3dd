#!/bin/bash
# round 2s: configs 3 / 5 and the sweep on the final library (one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02s_bench_cfg3.json 2>/dev/null; cut -c1-200 gpurun_out/r02s_bench_cfg3.json
timeout 600 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02s_bench_cfg5.json 2>/dev/null; cut -c1-200 gpurun_out/r02s_bench_cfg5.json
timeout 600 python tools/bench_sweep.py --samples 16 --repeats 4 > gpurun_out/r02s_sweep_n1.json 2>/dev/null; cut -c1-300 gpurun_out/r02s_sweep_n1.json
timeout 900 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r02s_train_b32.json 2>/dev/null; cut -c1-220 gpurun_out/r02s_train_b32.json
