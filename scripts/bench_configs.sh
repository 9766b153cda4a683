#!/bin/bash
# BASELINE.json configs other than the default one, same machinery (bench.py --config N): fills BASELINE.md section 4.
# usage (GPU box): bash scripts/bench_configs.sh [tag]
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
for cfg in 1 3 4 5; do
  echo "=== config $cfg"
  timeout 900 python bench.py --config $cfg --steps 7 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_config${cfg}.err | tee gpurun_out/${TAG}_bench_config${cfg}.json | cut -c1-900
  tail -2 gpurun_out/${TAG}_bench_config${cfg}.err
done
