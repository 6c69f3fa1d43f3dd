#!/bin/bash
# round 2o: recorded programs (native step replay), the pure-C host
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c_host.py tests/test_gpu_model.py tests/test_gpu_fp32_mode.py tests/test_gpu_vae.py tests/test_gpu_sweep.py -q 2>&1 | tail -25 > gpurun_out/r02o_tests.log; tail -12 gpurun_out/r02o_tests.log
for m in 1 0; do
  MFB_NATIVE_PROGRAM=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/r02o_bench_nograph_native$m.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02o_bench_nograph_native$m.json").read().strip().splitlines()[-1])
print("no graph, native program $m:", round(d["ms_per_step"],3), "ms/step  e2e", round(d["e2e"]["ms_per_step"],3))
PY
done
