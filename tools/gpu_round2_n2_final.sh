#!/bin/bash
# N = 2 at HEAD: default bench (ours + reference arm) under torchrun, as the driver launches them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02ad_bench_n2.json 2> gpurun_out/r02ad_bench_n2.err; tail -1 gpurun_out/r02ad_bench_n2.json | cut -c1-260; tail -4 gpurun_out/r02ad_bench_n2.err
