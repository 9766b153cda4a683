#!/bin/bash
# bench + ncu evidence.  Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
echo "=== bench"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
echo "=== ncu launch list of one eager step (timed region, --no-graph)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pww_timed/" -c 6000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
wc -l gpurun_out/launches.csv
echo "=== ncu full: xattn fwd + stats, B=2 and B=16"
for cfg in "2 1" "16 8"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_ -s 4 -c 2 -f \
      -o gpurun_out/xattn_B$1 python scripts/profile_xattn.py $1 $2 > gpurun_out/ncu_xattn_B$1.log 2>&1
  tail -1 gpurun_out/ncu_xattn_B$1.log
done
ls -la gpurun_out
