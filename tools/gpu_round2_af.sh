#!/bin/bash
# final-library evidence refresh: fine-tune step (config 4) bench line and the ncu launch list of one denoise step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 400 python bench.py --train --steps 6 --warmup 3 ) > gpurun_out/r02af_train_b32.json 2> gpurun_out/r02af_train_b32.err; tail -1 gpurun_out/r02af_train_b32.json | cut -c1-300; tail -3 gpurun_out/r02af_train_b32.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02af.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/bench_under_ncu_r02af.log 2>&1
wc -l gpurun_out/launches_r02af.csv
