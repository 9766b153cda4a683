#!/bin/bash
# round-2 call A: fused-kernel bring-up (each case isolated), full GPU tests, smoke, short bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
echo "=== bringup"
timeout 900 python scripts/fused_bringup.py 2>&1 | tee gpurun_out/fused_bringup.jsonl
echo "=== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
