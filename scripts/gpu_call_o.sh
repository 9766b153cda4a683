#!/bin/bash
# timelines of the final one-launch kernel (where does a job's time go?) + ncu of the two self-attention paths at N = 4096
set -u
mkdir -p gpurun_out
for cfg in "16 8 70" "16 0 70" "2 1 70" "2 1 5"; do
  set -- $cfg
  echo "=== timeline B=$1 biased=$2 cta=$3"
  timeout 300 python scripts/fused_timeline.py $1 $2 $3 2>&1 | tee gpurun_out/r02_final_timeline_B$1_b$2_cta$3.txt | head -70
done
echo "=== ncu self-attention N=4096 H=8 D=40: native vs cuDNN"
for impl in native sdpa; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc|fmha|flash|sdpa" -s 2 -c 1 -f \
      -o gpurun_out/r02_selfattn_$impl python scripts/profile_selfattn.py 2 4096 8 40 $impl > gpurun_out/r02_ncu_selfattn_$impl.log 2>&1
  tail -2 gpurun_out/r02_ncu_selfattn_$impl.log
done
