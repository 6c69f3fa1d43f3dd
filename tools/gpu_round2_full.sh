#!/bin/bash
# the driver's round-end sequence on one GPU: whole GPU suite, smoke(), default bench, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MFB_PARITY_LOG=gpurun_out/r02ae_parity_metrics.jsonl
rm -f $MFB_PARITY_LOG
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02ae_all_gpu_tests.log 2>&1; tail -6 gpurun_out/r02ae_all_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02ae_bench_n1.json 2> gpurun_out/r02ae_bench_n1.err; cut -c1-200 gpurun_out/r02ae_bench_n1.json; tail -4 gpurun_out/r02ae_bench_n1.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02ae_bench_reference.json 2> gpurun_out/r02ae_bench_reference.err; cut -c1-200 gpurun_out/r02ae_bench_reference.json; tail -4 gpurun_out/r02ae_bench_reference.err
