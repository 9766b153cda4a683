#!/bin/bash
# experiment: K / V tile copies with lanes along a row's chunks (few LSU wavefronts per instruction)
set -u
mkdir -p gpurun_out
for v in default kvcoal default kvcoal; do
  if [ $v = default ]; then unset PWW_B200_LIB; else export PWW_B200_LIB=$PWD/scripts/bin/libpww_$v.so; fi
  echo "=== variant $v: microbench N=4096 d=40"
  timeout 300 python scripts/xattn_microbench.py quick 2>&1 | tee -a gpurun_out/r02_kvcoal_${v}_microbench.jsonl
done
export PWW_B200_LIB=$PWD/scripts/bin/libpww_kvcoal.so
echo "=== kvcoal: tests"
timeout 600 python -m pytest tests/test_xattn_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "=== kvcoal: other shapes"
timeout 300 python scripts/xattn_microbench.py 2>&1 | tail -6 | tee gpurun_out/r02_kvcoal_microbench_sd15.jsonl
timeout 300 python scripts/xattn_microbench.py sd21 2>&1 | head -4 | tee gpurun_out/r02_kvcoal_microbench_sd21.jsonl
timeout 300 python scripts/fused_timeline.py 16 8 70 > gpurun_out/r02_kvcoal_timeline_B16_b8_cta70.txt 2>&1
unset PWW_B200_LIB
