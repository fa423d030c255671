#!/bin/bash
# Run every -m gpu test in its own process (a CUDA fault in one test cannot poison the others). Debug helper.
mkdir -p gpurun_out
out=gpurun_out/isolated.log
: > $out
for t in $(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::"); do
  CUDA_LAUNCH_BLOCKING=1 timeout 120 python -m pytest "$t" -x -q 2>&1 | tail -25 > gpurun_out/one.log
  if grep -q " passed" gpurun_out/one.log && ! grep -q "failed" gpurun_out/one.log; then echo "PASS $t" >> $out; else echo "FAIL $t" >> $out; grep -E "^E  " gpurun_out/one.log | head -12 >> $out; fi
done
cat $out
