#!/bin/bash
# round 2 (late): default bench with the in-graph family timing + e2e buffer swap; fresh ncu --set full of ten igemm launches of a step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02ab_bench_n1.json 2> gpurun_out/r02ab_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02ab_bench_n1.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(round(d["ms_per_step"],3), "ms/step  e2e", round(d["e2e"]["value"],3), "value", round(d["value"],3), d["clocks"])
print("eager families", r["families_ms_per_step"], "frac", round(r["frac"],4))
print("in_graph", r.get("in_graph"))
PY
tail -3 gpurun_out/r02ab_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 300 -c 10 -f -o gpurun_out/r02ab_igemm \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline --no-vae --no-report-dedup > gpurun_out/ncu_igemm_r02ab.log 2>&1
ls -la gpurun_out/r02ab_igemm.ncu-rep
