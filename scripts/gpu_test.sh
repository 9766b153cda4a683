#!/bin/bash
# quick GPU pass: parity tests (bounded).  usage: gpu_test.sh [pytest targets/args]
set -u
mkdir -p gpurun_out
T=${@:-tests}
timeout 900 python -m pytest $T -m gpu -x -q > gpurun_out/pytest_gpu_full.log 2>&1
grep -v "mbarrier timeout" gpurun_out/pytest_gpu_full.log | tail -40
grep "mbarrier timeout" gpurun_out/pytest_gpu_full.log | sort | uniq -c | sort -rn | head -30 > gpurun_out/timeouts.log
head -30 gpurun_out/timeouts.log
