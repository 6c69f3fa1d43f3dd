#!/bin/bash
# N = 8 check of the default bench (two launch streams) and the fine-tune step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02z_bench_n8.json 2> gpurun_out/r02z_bench_n8.err; tail -1 gpurun_out/r02z_bench_n8.json | cut -c1-220
timeout 500 $TR bench.py --gpus 8 --train --steps 6 --warmup 3 > gpurun_out/r02z_train_n8.json 2> gpurun_out/r02z_train_n8.err; tail -1 gpurun_out/r02z_train_n8.json | cut -c1-220
