#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
echo "=== 2-GPU sharding test"
timeout 600 python -m pytest tests/test_sharding_gpu.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_sharding_gpu.log
echo "=== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 27 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n2.err | tee gpurun_out/r02_bench_n2.json | cut -c1-1500
grep -i "nccl info.*nranks\|NCCL INFO comm" gpurun_out/bench_n2.err | head -4
