#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool (grouped-head one-launch kernel, N=1024 H=8 D=40 B=4, 2 biased, grid capped at 6 CTAs)"
  PWW_DEBUG_GRID=6 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_one.py 2>&1 | grep -E "SANITIZER|ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" | head -12 | tee gpurun_out/r02_sanitizer_$tool.txt
done
