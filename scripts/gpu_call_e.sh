#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "2 1 0" "16 8 0"; do
  timeout 120 python scripts/fused_timeline.py $cfg 2>&1 | tail -70 | tee gpurun_out/fused2_timeline_$(echo $cfg | tr ' ' '_').txt
done
