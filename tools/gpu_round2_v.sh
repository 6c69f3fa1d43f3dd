#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02x_all_gpu_tests.log 2>&1; tail -6 gpurun_out/r02x_all_gpu_tests.log
timeout 600 python tools/bench_sweep.py --samples 16 --repeats 4 > gpurun_out/r02x_sweep_n1.json 2>/dev/null; cut -c1-330 gpurun_out/r02x_sweep_n1.json
