#!/bin/bash
# Run each GPU test file in its own process (a faulting kernel must not poison the other files).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in tests/test_gpu_*.py; do
  echo "=== $f"
  timeout 900 python -m pytest "$f" -q -m gpu -x --no-header 2>&1 | tail -25
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
