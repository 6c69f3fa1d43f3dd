#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/bench_attn_bwd.py --batch 8 2>&1 | tee gpurun_out/r02f_attn_bwd_bench.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_dq_kernel|attn_bwd_dkv_kernel" -s 2 -c 2 -f -o gpurun_out/r02f_attn_bwd \
    python tools/bench_attn_bwd.py --batch 8 --only "self 64x64" --iters 1 > gpurun_out/r02f_ncu_attn_bwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"wgrad5_kernel" -s 20 -c 3 -f -o gpurun_out/r02f_wgrad5 \
    python tools/profile_train_step.py --batch 8 --steps 0 > gpurun_out/r02f_ncu_wgrad5.log 2>&1
timeout 300 python tools/bench_sweep.py --samples 16 --repeats 4 > gpurun_out/r02f_sweep_n1.json 2>/dev/null; cat gpurun_out/r02f_sweep_n1.json | cut -c1-400
ls -la gpurun_out/*.ncu-rep | tail -3
