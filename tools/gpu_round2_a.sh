#!/bin/bash
# Round 2, first GPU call: the new parity tests, the whole GPU suite, the default bench (with the eager-GPU baseline), configs 3 / 5,
# ncu --set full of the bandwidth kernels inside one real step, per-shape step profile, the fine-tune programs in bf16.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02a_new_tests.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_geometry.py --deselect tests/test_gpu_dropin.py 2>&1 | tail -15 | tee gpurun_out/r02a_all_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; tail -c 600 gpurun_out/r02a_bench_n1.err
timeout 300 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu-baseline --no-vae > gpurun_out/r02a_bench_cfg5.json 2> gpurun_out/r02a_bench_cfg5.err
timeout 300 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline --no-vae > gpurun_out/r02a_bench_cfg3.json 2> gpurun_out/r02a_bench_cfg3.err
cut -c1-400 gpurun_out/r02a_bench_n1.json; python - <<'PY'
import json
for f in ("n1","cfg5","cfg3"):
    try:
        d=json.loads(open(f"gpurun_out/r02a_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("gpu_eager_baseline"), d["roofline"]["families_ms_per_step"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
# ncu --set full of the bandwidth kernels inside one eager bench step (graph off so each is a plain launch)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_apply_kernel|gn_stats_kernel|gn_small_kernel|layernorm|cfg_sched_kernel" -s 181 -c 40 -f -o gpurun_out/r02a_norms \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-vae --no-report-dedup --no-eager-baseline > gpurun_out/r02a_ncu_norms.log 2>&1
timeout 300 python tools/profile_step.py > gpurun_out/r02a_profile_step.log 2>&1; head -8 gpurun_out/r02a_profile_step.log | cut -c1-600
timeout 120 python tools/bench_gn.py 2>&1 | tail -12 | tee gpurun_out/r02a_bench_gn.log
timeout 200 python tools/check_brushnet_trainer.py --precision bf16 --config tiny 2>&1 | tail -12 | tee gpurun_out/r02a_trainer_bf16_tiny.log
timeout 300 python tools/check_brushnet_trainer.py --precision bf16 --config sd15 --no-check --batch 8 --size 64 2>&1 | tail -4 | tee gpurun_out/r02a_trainer_bf16_sd15.log
ls -la gpurun_out | tail -12
