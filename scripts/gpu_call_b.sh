#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/fused_debug.py 2>&1 | tee gpurun_out/fused_debug.jsonl
echo "=== failing tests, full output"
timeout 600 python -m pytest tests/test_xattn_gpu.py -m gpu -q -k "aurora or long_job or all_biased or large_bias" 2>&1 | grep -E "^E  |assert|FAILED|passed|failed|Error" | cut -c1-300 | head -80 | tee gpurun_out/pytest_failing.log
echo "=== new pipeline tests"
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_pipeline.log
