#!/bin/bash
# experiment: Q producer warp of its own (TMA, ring of 3), output straight to global memory; variants: MUFU token, 4 heads per unit
set -u
mkdir -p gpurun_out
echo "=== smoke (default build)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for v in default xtoken g4 xtoken_g4; do
  if [ $v = default ]; then unset PWW_B200_LIB; else export PWW_B200_LIB=$PWD/scripts/bin/libpww_$v.so; fi
  echo "=== variant $v: microbench N=4096 d=40"
  timeout 300 python scripts/xattn_microbench.py quick 2>&1 | tee gpurun_out/r02_variant_${v}_microbench.jsonl
  echo "=== variant $v: tests"
  timeout 600 python -m pytest tests/test_xattn_gpu.py -m gpu -q -x -k "bias_path or long_job or job_table or key_lengths or stress" 2>&1 | tail -3
done
for v in default xtoken; do
  if [ $v = default ]; then unset PWW_B200_LIB; else export PWW_B200_LIB=$PWD/scripts/bin/libpww_$v.so; fi
  echo "=== variant $v: other head dims"
  timeout 300 python scripts/xattn_microbench.py 2>&1 | tail -6 | tee gpurun_out/r02_variant_${v}_microbench_sd15.jsonl
  timeout 300 python scripts/xattn_microbench.py sd21 2>&1 | head -4 | tee gpurun_out/r02_variant_${v}_microbench_sd21.jsonl
  for cfg in "16 8 70" "2 1 5"; do
    set -- $cfg
    timeout 300 python scripts/fused_timeline.py $1 $2 $3 > gpurun_out/r02_variant_${v}_timeline_B$1_b$2_cta$3.txt 2>&1
  done
done
unset PWW_B200_LIB
