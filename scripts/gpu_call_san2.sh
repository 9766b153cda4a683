#!/bin/bash
# compute-sanitizer memcheck + racecheck of the one-launch kernel at every head dim (ring mode: grid capped at 6 CTAs) and
# of the native self-attention kernel
set -u
mkdir -p gpurun_out
: > gpurun_out/r02_sanitizer_all.txt
for cfg in "1024 8 40 std self" "1152 5 64 max" "512 8 80 max" "256 8 160 std"; do
  for tool in memcheck racecheck; do
    echo "=== compute-sanitizer --tool $tool sanitize_one.py $cfg (B=4, 2 biased, grid capped at 6 CTAs)" | tee -a gpurun_out/r02_sanitizer_all.txt
    PWW_DEBUG_GRID=6 timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_one.py $cfg 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard|Invalid|Traceback" | head -12 | tee -a gpurun_out/r02_sanitizer_all.txt
  done
done
