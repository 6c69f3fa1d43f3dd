#!/bin/bash
# round 2 (late): single-pass cluster GroupNorm — op tests, per-shape microbench on/off, step bench on/off, parity at the goldens
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02gn}
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "groupnorm" -x 2>&1 | tail -15 | tee gpurun_out/${T}_gn_tests.txt
timeout 120 python tools/bench_gn.py > gpurun_out/${T}_bench_gn_cluster.txt 2>&1
MFB_GN_CLUSTER=0 timeout 120 python tools/bench_gn.py > gpurun_out/${T}_bench_gn_two_pass.txt 2>&1
paste gpurun_out/${T}_bench_gn_cluster.txt gpurun_out/${T}_bench_gn_two_pass.txt | cut -c1-200
run() {
  env $2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/${T}_$1.json 2>gpurun_out/${T}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$1.json").read().strip().splitlines()[-1])
    print("$1:", round(d["ms_per_step"],3), "ms/step", d["clocks"]["sm_mhz"], d["roofline"]["families_ms_per_step"], d["roofline"].get("groupnorm_hbm"), d["gpu_launches"])
except Exception as e:
    print("$1: ERR", e); print(open("gpurun_out/${T}_$1.err").read()[-2000:])
PY
}
run cluster_on MFB_GN_CLUSTER=1
run cluster_off MFB_GN_CLUSTER=0
run cluster_on_b MFB_GN_CLUSTER=1
run cluster_off_b MFB_GN_CLUSTER=0
MFB_PARITY_LOG=gpurun_out/${T}_parity_metrics.jsonl timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_geometry.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/${T}_model_tests.txt
