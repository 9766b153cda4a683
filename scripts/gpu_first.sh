#!/bin/bash
# First GPU pass: parity tests, smoke, short bench, ncu launch list.  Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
