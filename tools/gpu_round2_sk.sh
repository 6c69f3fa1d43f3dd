#!/bin/bash
# round 2 (late): split-K CTA pairs for the few-tile (8x8-level) igemm launches — op tests, per-shape microbench, step A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02sk}
timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "split_k or cluster_modes or block_n" -x 2>&1 | tail -25 | tee gpurun_out/${T}_tests.txt
grep -q "failed\|error\|Error" gpurun_out/${T}_tests.txt && exit 0
MFB_IGEMM_SPLITK=0 timeout 120 python tools/bench_igemm.py --only 8x8 2>&1 | tee gpurun_out/${T}_bench_igemm_off.txt
MFB_IGEMM_SPLITK=1 timeout 120 python tools/bench_igemm.py --only 8x8 2>&1 | tee gpurun_out/${T}_bench_igemm_on.txt
run() {
  env $2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/${T}_$1.json 2>gpurun_out/${T}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$1.json").read().strip().splitlines()[-1])
    print("$1:", round(d["ms_per_step"],3), "ms/step", d["clocks"]["sm_mhz"], d["roofline"]["families_ms_per_step"], d["roofline"]["in_graph"]["families_ms_per_step"])
except Exception as e:
    print("$1: ERR", e); print(open("gpurun_out/${T}_$1.err").read()[-2000:])
PY
}
run splitk_on MFB_IGEMM_SPLITK=1
run splitk_off MFB_IGEMM_SPLITK=0
run splitk_on_b MFB_IGEMM_SPLITK=1
run splitk_off_b MFB_IGEMM_SPLITK=0
