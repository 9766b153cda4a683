#!/bin/bash
# bench + ncu evidence.  Run under gpurun from the repo root.  usage: gpu_profile.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
echo "=== bench"
timeout 900 python bench.py --steps 27 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
echo "=== kernel micro-benchmarks"
timeout 600 python scripts/xattn_microbench.py dense > gpurun_out/${TAG}_xattn_microbench_sd15.jsonl 2>&1; cat gpurun_out/${TAG}_xattn_microbench_sd15.jsonl
timeout 600 python scripts/xattn_microbench.py sd21 dense > gpurun_out/${TAG}_xattn_microbench_sd21.jsonl 2>&1; cat gpurun_out/${TAG}_xattn_microbench_sd21.jsonl
echo "=== ncu launch list of one eager step (timed region, --no-graph)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pww_timed/" -c 6000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log
wc -l gpurun_out/${TAG}_launches.csv
echo "=== ncu full: one-launch xattn kernel, B=2 and B=16"
for cfg in "2 1" "16 8"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_fused -s 4 -c 1 -f \
      -o gpurun_out/${TAG}_xattn_B$1 python scripts/profile_xattn.py $1 $2 > gpurun_out/${TAG}_ncu_xattn_B$1.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_xattn_B$1.log
done
ls -la gpurun_out | tail -15
