# A/B of igemm library variants built by tools/ab_build.sh: parity tests + per-shape timing
for f in build/ab/*.so; do
  echo "=== $f"
  MFB200_LIB=$f timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "conv or linear or igemm or persistent or geglu or groupnorm_with" 2>&1 | tail -1
  MFB200_LIB=$f timeout 300 python tools/bench_igemm.py 2>&1 | grep -v "^$"
done
