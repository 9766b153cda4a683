#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in 5 1; do
  timeout 600 python bench.py --config $cfg --steps 7 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_bench_config${cfg}.err | tee gpurun_out/r02_bench_config${cfg}.json | cut -c1-300
done
