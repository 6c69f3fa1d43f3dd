#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q 2>&1 | tail -4
for f in "" "--two-streams"; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae $f > gpurun_out/r02v_bench$f.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02v_bench$f.json").read().strip().splitlines()[-1])
print("$f", round(d["ms_per_step"],3), "ms/step; dedup:", d.get("brushnet_cfg_dedup",{}).get("ms_per_step"), d["clocks"]["sm_mhz"])
PY
done
