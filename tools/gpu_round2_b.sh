#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_bf16.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02b_train_bf16_tests.log
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_zz_train_net.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02b_train_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vae --no-report-dedup > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 400 gpurun_out/r02b_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('gpu_eager_baseline'))"
