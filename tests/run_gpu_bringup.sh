#!/bin/bash
# Bring-up helper: runs every GPU test function in its own process (a trapped kernel poisons the CUDA context,
# so one failure must not hide the others) with a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
OUT=gpurun_out/bringup.log
: > $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv >> $OUT 2>&1
for f in "$@"; do
  for t in $(python -m pytest "$f" -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u); do
    echo "=== $t" >> $OUT
    timeout 300 python -m pytest "$t" -m gpu -x -q -s 2>&1 | tail -40 >> $OUT
    echo "exit: ${PIPESTATUS[0]}" >> $OUT
  done
done
grep -E "^=== |passed|failed|error|exit:" $OUT | tail -80
