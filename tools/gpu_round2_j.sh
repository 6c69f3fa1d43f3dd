#!/bin/bash
# round 2j: launch list of one fine-tune step at batch 32
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02j_train_launches_b32.csv python tools/profile_train_step.py --batch 32 --steps 1 > gpurun_out/r02j_train_ncu.log 2>&1; tail -2 gpurun_out/r02j_train_ncu.log
ls -la gpurun_out/r02j_train_launches_b32.csv
