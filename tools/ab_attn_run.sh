# A/B of the d=40 attention variants built by tools/ab_build.sh: parity tests + per-shape timing for each library
for f in build/ab/*.so; do
  echo "=== $f"
  MFB200_LIB=$f timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" 2>&1 | tail -1
  MFB200_LIB=$f timeout 100 python tools/bench_attn.py --only self 2>&1 | tail -4
done
