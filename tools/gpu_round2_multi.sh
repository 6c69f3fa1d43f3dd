#!/bin/bash
# Multi-GPU evidence for BASELINE configs 3 / 4 / 5 on N GPUs of one box:  tools/gpu_round2_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
RUNS=${2:-"cfg3 cfg5 cfg2 train sweep"}
has() { [[ " $RUNS " == *" $1 "* ]]; }
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-vae --no-report-dedup --no-eager-baseline"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
# config 3: 16 images / GPU (the sharded eval sweep's per-GPU load), NCCL rank log kept
has cfg3 && NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 $TR bench.py --gpus $N --config 3 $COMMON > gpurun_out/r02m_cfg3_n$N.json 2> gpurun_out/r02m_cfg3_n$N.err
has cfg3 && { grep -E "nranks|Connected all|NVLS" gpurun_out/r02m_cfg3_n$N.err | head -6 > gpurun_out/r02m_cfg3_n$N.nccl.txt; rm -f gpurun_out/r02m_cfg3_n$N.err; }
# config 5: 768x768 (9216-token attention), batch 4 / GPU
has cfg5 && timeout 600 $TR bench.py --gpus $N --config 5 $COMMON > gpurun_out/r02m_cfg5_n$N.json 2> /dev/null
# config 2 (headline workload) for reference on the same box
has cfg2 && timeout 600 $TR bench.py --gpus $N $COMMON > gpurun_out/r02m_cfg2_n$N.json 2> /dev/null
# config 4: fine-tune step, batch 32 / GPU, flat-gradient all-reduce over NCCL
has train && timeout 900 $TR bench.py --gpus $N --train --train-batch 32 --steps 4 --warmup 3 > gpurun_out/r02m_train_n$N.json 2> gpurun_out/r02m_train_n$N.err; has train && tail -c 600 gpurun_out/r02m_train_n$N.err
# the batched eval sweep end to end (uint8 in -> uint8 out), SHA-256 of all images must not depend on N
has sweep && timeout 900 $TR tools/bench_sweep.py --samples 16 --repeats 4 > gpurun_out/r02m_sweep_n$N.json 2> /dev/null
python - <<PY
import json
for f in ("cfg3","cfg5","cfg2","train","sweep"):
    try:
        d=json.loads(open(f"gpurun_out/r02m_{f}_n$N.json").read().strip().splitlines()[-1])
        print(f, "n=$N", d.get("ms_per_step"), d.get("value"), d.get("unit"), d.get("sha256_of_all_images","")[:16], (d.get("roofline") or {}).get("phases_ms"))
    except Exception as e: print(f, "ERR", e)
PY
cat gpurun_out/r02m_cfg3_n$N.nccl.txt 2>/dev/null; true
