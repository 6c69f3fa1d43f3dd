#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_zz_bench.py -q 2>&1 | tail -12
