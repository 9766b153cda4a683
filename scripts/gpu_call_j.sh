#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest xattn"
timeout 900 python -m pytest tests/test_xattn_gpu.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_xattn.log
echo "=== microbench N=4096"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r02_v2_microbench.jsonl
import json, sys, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
for (B, biased) in [(2, 1), (16, 8), (2, 0), (16, 0)]:
    r = bench.xattn_roofline(dev, B=B, biased=biased, iters=32 if B > 2 else 64)
    print(json.dumps({"B": B, "biased": biased, "us_op": round(r["us_op"], 2), "frac": round(r["alg_bytes"] / (r["us_op"] * 1e-6) / 1e9 / 6570, 3)}), flush=True)
PY
echo "=== launch floor"
timeout 120 scripts/bin/launch_floor 2>&1 | tee gpurun_out/r02_launch_floor.txt
echo "=== latent parity 512"
timeout 900 python scripts/latent_parity_512.py 2 2>&1 | tail -2 | tee gpurun_out/r02_latent_parity_512.json
