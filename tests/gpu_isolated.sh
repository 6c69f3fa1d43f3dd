#!/bin/bash
# Run every `-m gpu` test FUNCTION in its own process (a trapped kernel poisons the CUDA context of its process
# only).  usage: tests/gpu_isolated.sh [pytest path ...] -> summary on stdout, details in gpurun_out/gpu_isolated.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/gpu_isolated.log
: > "$LOG"
ids=$(python -m pytest --collect-only -q -m gpu "${@:-tests}" 2>/dev/null | grep '::' | sed 's/\[.*//' | sort -u)
pass=0; fail=0
for id in $ids; do
  if timeout 600 python -m pytest -q -m gpu "$id" >> "$LOG" 2>&1; then pass=$((pass+1)); echo "ok   $id"; else fail=$((fail+1)); echo "FAIL $id"; grep -E "^(FAILED|ERROR)|mfb200:|Error" "$LOG" | tail -15; fi
done
echo "isolated gpu test functions: $pass passed, $fail failed"
