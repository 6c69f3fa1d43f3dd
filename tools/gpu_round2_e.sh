#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_bf16.py tests/test_gpu_train.py tests/test_gpu_zz_train_net.py tests/test_gpu_zz_unet_train.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02e_train_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --train --train-batch 32 --steps 4 --warmup 3 > gpurun_out/r02e_train_b32.json 2> gpurun_out/r02e_train_b32.err; tail -c 800 gpurun_out/r02e_train_b32.err
python - <<'PY'
import json
for f in ("r02e_train_b32",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), d["value"], d["roofline"]["phases_ms"], d["roofline"]["achieved"], d["max_memory_gb"], d["clocks"])
    except Exception as e: print(f,"ERR",e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_train_launches.csv python tools/profile_train_step.py --batch 8 --steps 1 > gpurun_out/r02e_train_ncu.log 2>&1; tail -2 gpurun_out/r02e_train_ncu.log
