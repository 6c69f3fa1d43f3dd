#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_bench.py -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/r02zz_bench.json 2>gpurun_out/r02zz_bench.err; tail -2 gpurun_out/r02zz_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02zz_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["roofline"].get("attention_mufu"), d["config"]["streams"])
PY
