#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -k "attention or non_square" 2>&1 | tail -4
