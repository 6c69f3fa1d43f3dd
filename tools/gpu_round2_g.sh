#!/bin/bash
# round 2g: fp16 weights x bf16 activations (mixed-format tcgen05.mma), resize kernel, reference arm from baseline/_ref
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MFB_PARITY_LOG=gpurun_out/r02g_parity_metrics.jsonl
rm -f $MFB_PARITY_LOG
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "fp16 or linear or conv3x3 or geglu" 2>&1 | tail -15 > gpurun_out/r02g_ops_tests.log; tail -3 gpurun_out/r02g_ops_tests.log
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_geometry.py tests/test_gpu_sweep.py tests/test_gpu_vae.py tests/test_gpu_dropin.py -q 2>&1 | tail -40 > gpurun_out/r02g_model_tests.log; tail -5 gpurun_out/r02g_model_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; cut -c1-400 gpurun_out/r02g_bench_n1.json
export MFB_PARITY_LOG=gpurun_out/r02g_parity_metrics_gn_fused.jsonl
rm -f $MFB_PARITY_LOG
MFB_FUSE_GN_STATS=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_geometry.py -q 2>&1 | tail -15 > gpurun_out/r02g_model_tests_gn_fused.log; tail -3 gpurun_out/r02g_model_tests_gn_fused.log
MFB_FUSE_GN_STATS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae > gpurun_out/r02g_bench_n1_gn_fused.json 2>/dev/null; cut -c1-300 gpurun_out/r02g_bench_n1_gn_fused.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02g_bench_reference.json 2> gpurun_out/r02g_bench_reference.err; cut -c1-600 gpurun_out/r02g_bench_reference.json
