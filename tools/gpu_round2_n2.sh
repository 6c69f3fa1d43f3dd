#!/bin/bash
# N = 2 sanity after this round's changes: inference bench, reference arm under torchrun, fine-tune step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err; tail -1 gpurun_out/r02p_bench_n2.json | cut -c1-220
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02p_bench_reference_n2.json 2> gpurun_out/r02p_bench_reference_n2.err; tail -1 gpurun_out/r02p_bench_reference_n2.json | cut -c1-220
timeout 900 $TR bench.py --gpus 2 --train --steps 6 --warmup 3 > gpurun_out/r02p_train_n2.json 2> gpurun_out/r02p_train_n2.err; tail -1 gpurun_out/r02p_train_n2.json | cut -c1-260; tail -2 gpurun_out/r02p_train_n2.err
