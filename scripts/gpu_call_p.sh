#!/bin/bash
# experiment: Q / map by cp.async from a loader warp of its own (no TMA), ring of 3, output straight to global memory
set -u
mkdir -p gpurun_out
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== microbench sd15"
timeout 600 python scripts/xattn_microbench.py 2>&1 | tee gpurun_out/r02_microbench_sd15_cpasync.jsonl
echo "=== microbench sd21"
timeout 600 python scripts/xattn_microbench.py sd21 2>&1 | tee gpurun_out/r02_microbench_sd21_cpasync.jsonl
for cfg in "16 8 70" "2 1 5"; do
  set -- $cfg
  echo "=== timeline B=$1 biased=$2 cta=$3"
  timeout 300 python scripts/fused_timeline.py $1 $2 $3 2>&1 | tee gpurun_out/r02_cpasync_timeline_B$1_b$2_cta$3.txt | head -52
done
echo "=== pytest xattn + schedule"
timeout 900 python -m pytest tests/test_xattn_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_xattn_cpasync.log
echo "=== ncu self-attention native"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 2 -c 1 -f \
    -o gpurun_out/r02_selfattn_native python scripts/profile_selfattn.py 2 4096 8 40 native > gpurun_out/r02_ncu_selfattn_native.log 2>&1
tail -2 gpurun_out/r02_ncu_selfattn_native.log
