#!/bin/bash
# One GPU call that refreshes the judged evidence: launch list of one bench step, ncu --set full of the dominant kernels,
# per-shape step profile.  usage: tools/gpu_profile_round.sh <tag>   (outputs under gpurun_out/, summaries made locally)
tag=${1:-r01x}
cd "$(dirname "$0")/.."
# launch list: graph capture off so every kernel is a plain launch; skip the warm-up pass, keep one step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/bench_under_ncu_${tag}.log 2>&1
# full captures: the d=40 self-attention (one launch) and ten consecutive igemm launches of a step
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 4 -c 1 -f -o gpurun_out/${tag}_attn40 \
    python tools/bench_attn.py --only "self 64x64" > gpurun_out/ncu_attn_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 300 -c 10 -f -o gpurun_out/${tag}_igemm \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/ncu_igemm_${tag}.log 2>&1
timeout 300 python tools/profile_step.py > gpurun_out/profile_step_${tag}.log 2>&1
tail -3 gpurun_out/profile_step_${tag}.log
ls -la gpurun_out | tail -8
