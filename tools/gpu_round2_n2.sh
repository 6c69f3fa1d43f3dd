#!/bin/bash
# N = 2 sanity with the two-stream default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02w_bench_n2.json 2> gpurun_out/r02w_bench_n2.err; tail -1 gpurun_out/r02w_bench_n2.json | cut -c1-220
timeout 600 $TR bench.py --gpus 2 --config 3 --steps 20 --warmup 5 > gpurun_out/r02w_cfg3_n2.json 2>/dev/null; tail -1 gpurun_out/r02w_cfg3_n2.json | cut -c1-220
timeout 600 $TR bench.py --gpus 2 --config 5 --steps 20 --warmup 5 > gpurun_out/r02w_cfg5_n2.json 2>/dev/null; tail -1 gpurun_out/r02w_cfg5_n2.json | cut -c1-220
