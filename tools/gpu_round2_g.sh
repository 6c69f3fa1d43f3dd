#!/bin/bash
# round 2g: resize kernel, reference arm from baseline/_ref (CPU + eager bf16 on the B200), GN-stats fusion A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MFB_PARITY_LOG=gpurun_out/r02g_parity_metrics.jsonl
rm -f $MFB_PARITY_LOG
timeout 1200 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_model.py tests/test_gpu_geometry.py -q 2>&1 | tail -40 > gpurun_out/r02g_model_tests.log; tail -5 gpurun_out/r02g_model_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; cut -c1-400 gpurun_out/r02g_bench_n1.json
timeout 900 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02g_bench_cfg5.json 2> gpurun_out/r02g_bench_cfg5.err; cut -c1-300 gpurun_out/r02g_bench_cfg5.json
