#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_zz_train_net.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02d_train_tests.log
timeout 600 python -m pytest tests/test_gpu_zz_unet_train.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02d_unet_train_tests.log
timeout 900 python bench.py --train --train-batch 32 --steps 4 --warmup 3 > gpurun_out/r02d_train_b32.json 2> gpurun_out/r02d_train_b32.err; tail -c 800 gpurun_out/r02d_train_b32.err
MFB_WGRAD_LEGACY=1 timeout 900 python bench.py --train --train-batch 32 --steps 4 --warmup 3 > gpurun_out/r02d_train_b32_legacy.json 2> gpurun_out/r02d_train_b32_legacy.err
python - <<'PY'
import json
for f in ("r02d_train_b32","r02d_train_b32_legacy"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), d["value"], d["roofline"]["phases_ms"], d["roofline"]["achieved"], d["max_memory_gb"], d["clocks"])
    except Exception as e: print(f,"ERR",e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_train_launches.csv python tools/profile_train_step.py --batch 8 --steps 1 > gpurun_out/r02d_train_ncu.log 2>&1; tail -2 gpurun_out/r02d_train_ncu.log
ls -la gpurun_out/r02d_train_launches.csv
