#!/bin/bash
# last full validation of the round: every GPU test, smoke, the default bench line
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default flags)"
timeout 1200 python bench.py 2> gpurun_out/bench_final.err | tee gpurun_out/r02_bench_n1_final.json | cut -c1-1500
