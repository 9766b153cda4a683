#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest xattn"
timeout 1200 python -m pytest tests/test_xattn_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_xattn.log
echo "=== timelines"
for cfg in "2 1 0" "2 1 1" "2 1 126" "16 8 0"; do
  timeout 120 python scripts/fused_timeline.py $cfg 2>&1 | tail -70 | tee gpurun_out/fused2_timeline_$(echo $cfg | tr ' ' '_').txt
done
echo "=== microbench"
timeout 600 python scripts/xattn_microbench.py 2>&1 | tee gpurun_out/r02_v2_microbench.jsonl
