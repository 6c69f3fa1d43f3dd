#!/bin/bash
# round 2u: BrushNet || UNet on two launch streams once the conv_in-site tap no longer serialises them; capped persistent igemm grids
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -k "two_stream or fused or golden or dedup or guess" 2>&1 | tail -3
run() {   # label, env cap, extra flags
  MFB_IGEMM_MAX_CTAS=$2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup $3 > gpurun_out/r02u_$1.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02u_$1.json").read().strip().splitlines()[-1])
print("$1:", round(d["ms_per_step"],3), "ms/step", d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
PY
}
run base_148 0 ""
run two_148 0 "--two-streams"
run two_132 132 "--two-streams"
run two_111 111 "--two-streams"
run two_74 74 "--two-streams"
run base_148b 0 ""
run two_148b 0 "--two-streams"
