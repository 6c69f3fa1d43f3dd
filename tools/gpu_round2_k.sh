#!/bin/bash
# round 2k: A/B of cta_group::2 pair mode on the few-tile (8x8 level) launches only
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 0 2 1; do
  echo "== MFB_IGEMM_MODE_SMALLM=$m"
  MFB_IGEMM_MODE_SMALLM=$m timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "8x8 or cluster or conv3x3" 2>&1 | tail -2
  MFB_IGEMM_MODE_SMALLM=$m timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/r02k_bench_smallm$m.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02k_bench_smallm$m.json").read().strip().splitlines()[-1])
print("mode $m", round(d["ms_per_step"],3), d["roofline"]["families_ms_per_step"], d["clocks"]["sm_mhz"])
PY
done
MFB_IGEMM_MODE_SMALLM=2 timeout 600 python tools/profile_step.py 2>/dev/null | grep "M=1024" | head -20
