#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== microbench sd15 (fused vs dense pair)"
timeout 900 python scripts/xattn_microbench.py dense 2>&1 | tee gpurun_out/r02_microbench_sd15.jsonl
echo "=== microbench sd21"
timeout 900 python scripts/xattn_microbench.py sd21 dense 2>&1 | tee gpurun_out/r02_microbench_sd21.jsonl
echo "=== self-attention microbench"
timeout 600 python scripts/selfattn_microbench.py 2>&1 | tee gpurun_out/r02_selfattn_microbench.jsonl
