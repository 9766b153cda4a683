#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02_sanitizer_d40.txt
for cfg in "1024 8 40 max" "1024 8 40 std"; do
  echo "=== compute-sanitizer --tool racecheck sanitize_one.py $cfg (B=4, 2 biased, grid capped at 6 CTAs)" | tee -a gpurun_out/r02_sanitizer_d40.txt
  PWW_DEBUG_GRID=6 timeout 300 compute-sanitizer --tool racecheck --print-limit 4 python scripts/sanitize_one.py $cfg 2>&1 | grep -E "RACECHECK SUMMARY|^ok|Error|hazard|Traceback" | head -8 | tee -a gpurun_out/r02_sanitizer_d40.txt
done
