#!/bin/bash
# Build kernel A/B variants of libmfb200.so into build/ab/<name>.so (git-ignored; they travel with gpurun).
# usage: tools/ab_build.sh name "-DMACRO=.. -DMACRO2=.." [name2 "flags2" ...]; select one with MFB200_LIB=build/ab/<name>.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/reflecting-reality_b200/csrc
mkdir -p $ROOT/build/ab
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
       -o $ROOT/build/ab/$name.so $CSRC/api.cu $CSRC/igemm.cu $CSRC/norm.cu $CSRC/misc.cu $CSRC/attn.cu $CSRC/fp32mode.cu $CSRC/train.cu &
done
wait
ls -la $ROOT/build/ab
