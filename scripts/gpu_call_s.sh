#!/bin/bash
# experiment: Q ring of 3 at head dim 40 (exchange area halved to make room), with and without the MUFU token
set -u
mkdir -p gpurun_out
for v in default nq3 nq3_xtoken default nq3 nq3_xtoken; do
  if [ $v = default ]; then unset PWW_B200_LIB; else export PWW_B200_LIB=$PWD/scripts/bin/libpww_$v.so; fi
  echo "=== variant $v: microbench N=4096 d=40"
  timeout 300 python scripts/xattn_microbench.py quick 2>&1 | tee -a gpurun_out/r02_nq3_${v}_microbench.jsonl
done
for v in nq3 nq3_xtoken; do
  export PWW_B200_LIB=$PWD/scripts/bin/libpww_$v.so
  echo "=== $v: tests"
  timeout 600 python -m pytest tests/test_xattn_gpu.py -m gpu -q -x 2>&1 | tail -3
  timeout 300 python scripts/fused_timeline.py 16 8 70 > gpurun_out/r02_${v}_timeline_B16_b8_cta70.txt 2>&1
done
unset PWW_B200_LIB
