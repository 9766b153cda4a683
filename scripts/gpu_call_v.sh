#!/bin/bash
timeout 120 python -m pytest tests/test_xattn_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_last.log
