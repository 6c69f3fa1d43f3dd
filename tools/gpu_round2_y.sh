#!/bin/bash
# round 2y: programmatic dependent launch (MFB_PDL=1) re-measured with the two-stream schedule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  MFB_PDL=$2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup $3 > gpurun_out/r02y_$1.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02y_$1.json").read().strip().splitlines()[-1])
print("$1:", round(d["ms_per_step"],3), "ms/step", d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
PY
}
run two_pdl0 0 ""
run two_pdl1 1 ""
run one_pdl0 0 "--no-two-streams"
run one_pdl1 1 "--no-two-streams"
run two_pdl0b 0 ""
run two_pdl1b 1 ""
