#!/bin/bash
# Runs every `-m gpu` test in its own process (a CUDA fault in one test cannot poison the others) and writes
# a summary to gpurun_out/.  Usage (on the GPU box, from the repo root):  bash tools/gpu_tests_isolated.sh [-k expr]
mkdir -p gpurun_out
OUT=gpurun_out/gpu_tests_isolated.log
: > "$OUT"
ids=$(python -m pytest tests -m gpu --collect-only -q "$@" 2>/dev/null | grep '::')
pass=0; fail=0
for id in $ids; do
  echo "=== $id" >> "$OUT"
  if timeout 600 python -m pytest "$id" -x -q -p no:cacheprovider >> "$OUT" 2>&1; then pass=$((pass+1)); echo "PASS $id"; else fail=$((fail+1)); echo "FAIL $id"; fi
done
echo "isolated gpu tests: $pass passed, $fail failed" | tee -a "$OUT"
[ "$fail" -eq 0 ]
