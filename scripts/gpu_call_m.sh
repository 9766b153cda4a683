#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest xattn"
timeout 900 python -m pytest tests/test_xattn_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_xattn.log
echo "=== microbench sd21 (fused vs dense pair)"
timeout 900 python scripts/xattn_microbench.py sd21 dense 2>&1 | tee gpurun_out/r02_microbench_sd21_final.jsonl
