#!/bin/bash
# round 2y2: the bit-identical igemm knobs re-measured under the two-stream schedule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  env $2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/r02y2_$1.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02y2_$1.json").read().strip().splitlines()[-1])
print("$1:", round(d["ms_per_step"],3), "ms/step", d["clocks"]["sm_mhz"], d["roofline"]["families_ms_per_step"])
PY
}
run base A=0
run epiw1 MFB_IGEMM_EPIW=1
run epiw16 MFB_IGEMM_EPIW=16
run nstg640 MFB_IGEMM_NSTG_KMAX=640
run base_b A=0
run epiw1_b MFB_IGEMM_EPIW=1
