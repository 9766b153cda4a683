#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "=== timelines"
for cfg in "2 1 0" "2 1 75" "2 1 147" "16 8 0" "16 8 75"; do
  timeout 120 python scripts/fused_timeline.py $cfg 2>&1 | tail -60 | tee gpurun_out/fused_timeline_$(echo $cfg | tr ' ' '_').txt
done
echo "=== ncu full"
for cfg in "2 1" "16 8"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_fused -s 4 -c 1 -f \
      -o gpurun_out/r02_xattn_B$1 python scripts/profile_xattn.py $1 $2 4096 8 40 6 > gpurun_out/r02_ncu_xattn_B$1.log 2>&1
  tail -1 gpurun_out/r02_ncu_xattn_B$1.log
done
ls -la gpurun_out | tail -12
