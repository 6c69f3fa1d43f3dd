#!/bin/bash
# round 2t: same-box A/B of the kept GroupNorm statistics in the fine-tune step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 1 0 1 0; do
  MFB_TRAIN_KEEP_GN_STATS=$k timeout 900 python bench.py --train --steps 10 --warmup 3 > gpurun_out/r02t_train_keep$k.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02t_train_keep$k.json").read().strip().splitlines()[-1])
print("keep stats $k:", round(d["ms_per_step"],2), "ms/step", d["roofline"]["phases_ms"], d["clocks"]["sm_mhz"])
PY
done
