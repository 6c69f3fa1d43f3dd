#!/bin/bash
# round 2y3: priority of the BrushNet launch stream
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  env $2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/r02y3_$1.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02y3_$1.json").read().strip().splitlines()[-1])
print("$1:", round(d["ms_per_step"],3), "ms/step", d["clocks"]["sm_mhz"])
PY
}
run prio0 MFB_SIDE_PRIORITY=0
run prio_hi MFB_SIDE_PRIORITY=-1
run prio0_b MFB_SIDE_PRIORITY=0
run prio_hi_b MFB_SIDE_PRIORITY=-1
