#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== configs 3 and 5"
for cfg in 3 5; do
  timeout 900 python bench.py --config $cfg --steps 7 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_bench_config${cfg}.err | tee gpurun_out/r02_bench_config${cfg}.json | cut -c1-400
done
echo "=== step kineto"
timeout 600 python scripts/step_profile.py 2>&1 | tail -60 > gpurun_out/r02_step_kineto.txt; tail -12 gpurun_out/r02_step_kineto.txt | cut -c1-160
echo "=== ncu launch list of one eager step (timed region, --no-graph)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pww_timed/" -c 6000 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-graph > gpurun_out/r02_ncu_bench.log 2>&1
wc -l gpurun_out/r02_launches.csv
echo "=== ncu full: one-launch kernel, B=2 and B=16"
for cfg in "2 1" "16 8"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_fused2 -s 4 -c 1 -f \
      -o gpurun_out/r02_xattn_B$1 python scripts/profile_xattn.py $1 $2 4096 8 40 6 > gpurun_out/r02_ncu_xattn_B$1.log 2>&1
  tail -1 gpurun_out/r02_ncu_xattn_B$1.log
done
