#!/bin/bash
# ncu --set full of one 8x8-level conv launch: split-K pair (MODE 3) against the shipped 128 x 80 tiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 0; do
  MFB_IGEMM_SPLITK=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 3 -c 1 -f -o gpurun_out/r02sk_ncu_splitk$v \
      python tools/bench_igemm.py --only "8x8 1280->1280" --iters 2 > gpurun_out/r02sk_ncu_splitk$v.log 2>&1
  tail -2 gpurun_out/r02sk_ncu_splitk$v.log
done
ls -la gpurun_out/r02sk_ncu_*.ncu-rep
