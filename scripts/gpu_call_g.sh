#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest selfattn"
timeout 600 python -m pytest tests/test_selfattn_gpu.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_selfattn.log
echo "=== stress"
timeout 900 python scripts/selfattn_stress.py 2>&1 | cut -c1-1500 | tee gpurun_out/selfattn_stress.jsonl
echo "=== self-attention microbench"
timeout 600 python scripts/selfattn_microbench.py 2>&1 | tee gpurun_out/r02_selfattn_microbench.jsonl
