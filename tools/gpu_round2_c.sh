#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_bf16.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02c_train_bf16_tests.log
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_zz_train_net.py tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02c_train_tests.log
timeout 600 python -m pytest tests/test_gpu_zz_unet_train.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02c_unet_train_tests.log
MFB_FUSE_GN_STATS=1 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_geometry.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02c_fused_stats_model_tests.log
cp gpurun_out/parity_metrics.jsonl gpurun_out/r02c_parity_fused_stats.jsonl 2>/dev/null
for f in 0 1; do
  MFB_FUSE_GN_STATS=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vae --no-report-dedup --no-eager-baseline > gpurun_out/r02c_bench_fuse$f.json 2> gpurun_out/r02c_bench_fuse$f.err
done
python - <<'PY'
import json
for f in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/r02c_bench_fuse{f}.json").read().strip().splitlines()[-1])
        print("fuse",f, round(d["ms_per_step"],3), d["roofline"]["families_ms_per_step"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 python bench.py --train --train-batch 8 --steps 3 --warmup 2 > gpurun_out/r02c_train_b8.json 2> gpurun_out/r02c_train_b8.err; tail -c 1500 gpurun_out/r02c_train_b8.err; cut -c1-1500 gpurun_out/r02c_train_b8.json
timeout 900 python bench.py --train --train-batch 32 --steps 3 --warmup 2 > gpurun_out/r02c_train_b32.json 2> gpurun_out/r02c_train_b32.err; tail -c 1500 gpurun_out/r02c_train_b32.err; cut -c1-1500 gpurun_out/r02c_train_b32.json
